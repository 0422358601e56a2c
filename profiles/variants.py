"""Times bench.py (single GPU, --no-cpu) against each prebuilt library variant under
particles.jl_b200/_variants/ (compile-time tuning knobs) and prints one line per variant."""
import glob, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [("base", None)] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(root, "particles.jl_b200/_variants/lib_*.so")))]
for name, path in libs:
    env = dict(os.environ)
    if path: env["PARTICLES_B200_LIB"] = path
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu"] + sys.argv[1:], env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        pk = j["roofline"]["per_kernel"]
        line = {"variant": name, "ms_per_step": round(j["ms_per_step"], 4), "e2e": j["e2e"]["value"],
                **{k: round(v["share"] * j["ms_per_step"], 4) for k, v in pk.items()}}
    except Exception as e:
        line = {"variant": name, "error": str(e), "stderr": r.stderr[-300:]}
    print(json.dumps(line), flush=True)
