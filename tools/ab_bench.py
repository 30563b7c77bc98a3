"""A/B of library builds on the GPU box: python tools/ab_bench.py [--workload W] lib1.so lib2.so ...
Runs bench.py (device-resident part only) once per library (SACX_LIB) and prints updates/s and ms per update."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
wl = "bipedal"
if args and args[0] == "--workload":
    wl, args = args[1], args[2:]
for lib in args:
    env = dict(os.environ, SACX_LIB=os.path.abspath(lib))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--no-extras", "--no-cpu-baseline", "--steps", "2000" if wl in ("bipedal", "donkey") else "20",
                        "--warmup", "100" if wl in ("bipedal", "donkey") else "5", "--e2e-steps", "400"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f"{os.path.basename(lib):24s} {wl}: {d['value']:10.1f} /s  {d['ms_per_step'] * 1000:8.2f} us/step  e2e {e.get('value', float('nan')):9.1f}/s  launches {d['gpu_launches']}")
    except Exception as ex:
        print(lib, "FAILED", ex, p.stderr[-800:])
