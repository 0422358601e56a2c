"""A/B of run-time and build-time knobs in ONE gpurun call: each spec `name[:ENV=V,...][@variant]` runs
`bench.py --no-cpu` (single GPU) with the environment / prebuilt library variant it names
(particles.jl_b200/_variants/lib_<variant>.so) and prints one line: ms per observation, per-class ms, evidence crc.
Usage: python profiles/ab_env.py [--bench-args "..."] spec [spec ...]"""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
bench_args = []
if args and args[0] == "--bench-args":
    bench_args = args[1].split(); args = args[2:]
for spec in args:
    lib = None
    if "@" in spec: spec, lib = spec.split("@", 1)
    name, _, envs = spec.partition(":")
    env = dict(os.environ)
    for kv in filter(None, envs.split(",")):
        k, _, v = kv.partition("="); env[k] = v
    if lib: env["PARTICLES_B200_LIB"] = os.path.join(root, "particles.jl_b200/_variants/lib_%s.so" % lib)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu"] + bench_args, env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        pk = j["roofline"]["per_kernel"]
        line = {"variant": name + ("@" + lib if lib else ""), "ms_per_step": round(j["ms_per_step"], 4),
                "e2e_ms": round(j["e2e"]["ms_per_step"], 4), "crc": j["config"]["evidence_check"]["crc32"],
                **{k: round(v["ms_per_step"], 4) for k, v in pk.items()}}
    except Exception as e:
        line = {"variant": name, "error": str(e), "stderr": r.stderr[-400:]}
    print(json.dumps(line), flush=True)
