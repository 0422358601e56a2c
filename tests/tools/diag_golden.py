"""Diagnostic: first step at which the device (sorted order) and the oracle disagree on a golden fixture."""
import os, sys
import numpy as np
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(R, "particles.jl_b200")); sys.path.insert(0, R)
import particles_b200 as pb
from oracle import oracle as orc

g = np.load(os.path.join(R, "tests", "golden", (sys.argv[1] if len(sys.argv) > 1 else "fearnhead_niw_2d_n500") + ".npz"))
N, T = int(g["N"]), len(g["ys"])
dev = pb.FearnheadParticles(N, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0),
                            history=T, order=pb.PF_ORDER_SORTED)
ref = orc.OracleFilter(orc.FEARNHEAD, N, orc.Prior.niw(np.zeros(2), 0.1, np.eye(2), 3.0), 1.0, order=orc.ORDER_SORTED)
for t, (y, u) in enumerate(zip(g["ys"], g["uniforms"])):
    ref.fh_expand(y)
    w_ref = ref.putative_weights().copy()
    par = ref.putative_parents()[0].copy()
    dev.fit_(y, [u])
    ref.fh_step(y, u)
    a_dev, a_ref = dev.last_ancestors(), ref.last_ancestors()
    if len(a_dev) == len(a_ref) and np.array_equal(a_dev, a_ref) and np.array_equal(dev.last_assignments() + 1, ref.last_assignments()):
        continue
    w_dev = dev.putative_weights()
    s = dev.last_step()
    print("first mismatch at step", t, "M", len(w_ref), len(w_dev), "n_out", len(a_dev), len(a_ref),
          "kept", s["n_kept"], ref.stat("n_kept"), "resamp", s["n_resamp_got"], ref.stat("n_resamp_got"), "u", u)
    if len(w_ref) == len(w_dev):
        rel = np.abs(w_dev - w_ref) / np.maximum(w_ref, 1e-300)
        print("max rel weight diff", rel.max(), "at", rel.argmax())
        o_dev, o_ref = np.argsort(w_dev, kind="stable"), np.argsort(w_ref, kind="stable")
        nd = np.nonzero(o_dev != o_ref)[0]
        print("sorted order differs at", len(nd), "positions", nd[:10])
        for p in nd[:6]:
            i, j = o_dev[p], o_ref[p]
            print("  pos", p, "dev item", i, w_dev[i].hex(), w_ref[i].hex(), "| ref item", j, w_dev[j].hex(), w_ref[j].hex())
        _, inv_d = np.unique(w_dev, return_inverse=True); _, inv_r = np.unique(w_ref, return_inverse=True)
        print("tie structure equal:", np.array_equal(inv_d, inv_r), "groups", inv_d.max() + 1, inv_r.max() + 1)
        for name, w in (("dev weights", w_dev), ("ref weights", w_ref)):
            sl, fl, c, nk = orc.select_literal(w, N, u, orc.ORDER_SORTED)
            se, fe, Tn, nke = orc.select_exact(w, N, u, orc.ORDER_SORTED)
            print(name, ": literal==exact", np.array_equal(sl, se), "| literal parents==dev", np.array_equal(par[sl], a_dev),
                  "| exact parents==dev", np.array_equal(par[se], a_dev), "| literal parents==ref", np.array_equal(par[sl], a_ref),
                  "| exact parents==ref", np.array_equal(par[se], a_ref))
    d = np.nonzero(a_dev[:min(len(a_dev), len(a_ref))] != a_ref[:min(len(a_dev), len(a_ref))])[0]
    print("ancestor positions differing", len(d), d[:10], a_dev[d[:10]], a_ref[d[:10]])
    break
else:
    print("no mismatch")
