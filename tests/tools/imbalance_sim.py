"""CPU estimate (oracle, index order) of the producer-side imbalance of the sharded tail: with the
population re-partitioned into G equal blocks every step, how many survivors does each block
produce in the next step?  max over ranks / mean is the skew the exchange barrier waits for."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle as orc
from bench import mixture

G = 8
for log2n in (12, 14, 16):
    N, T = 1 << log2n, 140
    rng = np.random.default_rng(100)
    ys = mixture(rng, T); us = rng.random(T)
    f = orc.OracleFilter(orc.FEARNHEAD, N, orc.Prior.niw(np.zeros(2), 0.1, np.eye(2), 3.0), 1.0, order=orc.ORDER_INDEX, keep_history=False)
    ratios = []
    for t in range(T):
        n_before = f.n
        f.fh_step(ys[t], us[t])
        if n_before == N and f.n == N and t >= 60:
            anc = f.last_ancestors()
            blk = (anc.astype(np.int64) * G) // N
            cnt = np.bincount(blk, minlength=G)
            ratios.append(cnt.max() / cnt.mean())
    r = np.array(ratios)
    print(f"N=2^{log2n}: steps {len(r)}  max/mean survivors per block: median {np.median(r):.3f}  p90 {np.quantile(r, .9):.3f}  max {r.max():.3f}")
