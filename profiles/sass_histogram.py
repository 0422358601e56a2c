#!/usr/bin/env python
"""Static SASS opcode histogram of the hot kernels in the built library (cuobjdump -sass): what the instruction stream
is made of, and which Blackwell-specific instructions are (not) in it.  Writes profiles/r2/sass_histograms.json."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "particles.jl_b200", "libparticles_b200.so")
KERNELS = {
    "k_expand<2,2,INST,double> (instantiate + next expansion)": r"k_expandILi2ELi2ELb1EdLb0ELb0EE",
    "k_expand<2,2,INST,float> (PF_F32)": r"k_expandILi2ELi2ELb1EfLb0ELb0EE",
    "k_expand<2,2,INST,double,REMOTE> (sharded: children delivered to their new owners)": r"k_expandILi2ELi2ELb1EdLb1ELb0EE",
    "k_expand<2,2,false,double> (expansion from the store)": r"k_expandILi2ELi2ELb0EdLb0ELb0EE",
    "k_instantiate<2,double>": r"k_instantiateILi2EdE",
    "k_scan_apply": r"k_scan_apply",
    "k_sel_bracket": r"k_sel_bracket",
    "k_sel_solve": r"k_sel_solve",
    "k_cl_propagate<8,double>": r"k_cl_propagateILi8EdE",
    "k_cl_finish<8,double>": r"k_cl_finishILi8EdE",
    "k_similarity": r"k_similarity",
}


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    res = {}
    for label, pat in KERNELS.items():
        body = next((f for f in funcs if re.match(r"_Z\d+" + pat, f) or re.match(pat, f)), None)
        if body is None:
            res[label] = None
            continue
        ops = collections.Counter()
        for line in body.splitlines():
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
        n = sum(ops.values())
        res[label] = {"instructions": n, "top": {k: v for k, v in ops.most_common(14)},
                      "local_memory": {k: ops[k] for k in ("LDL", "STL") if ops[k]},
                      "fp64": sum(v for k, v in ops.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "MUFU")),
                      "async_copy": {k: ops[k] for k in ("LDGSTS", "LDGDEPBAR", "DEPBAR") if ops[k]},
                      "blackwell_specific": {k: ops[k] for k in ("UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "SYNCS", "UBLKCP") if ops[k]}}
    os.makedirs(os.path.join(ROOT, "profiles", "r2"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "profiles", "r2", "sass_histograms.json"), "w"), indent=1)
    for k, v in res.items():
        print(k, "->", "missing" if v is None else (v["instructions"], v["local_memory"], v["blackwell_specific"], list(v["top"].items())[:6]))


if __name__ == "__main__":
    sys.exit(main())
