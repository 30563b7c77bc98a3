"""profiles/sass_summary.txt: mnemonic counts per kernel of the shipped library (cuobjdump -sass), so that the tcgen05 / TMA / TMEM
claims can be checked without disassembling it again.   python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "soft-actor-critic_b200", "lib", "libsacx.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "HMMA", "FFMA", "SYNCS", "LDGSTS", "UBLKCP", "ATOM", "RED",
        "BAR.SYNC", "MEMBAR", "LDS", "STS", "LDG", "STG"]
fn, cnt = None, collections.defaultdict(collections.Counter)
for L in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", L)
    if m:
        fn = m.group(1)
        continue
    if fn is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", L):
        continue
    cnt[fn]["instructions"] += 1
    for p in pats:
        if re.search(r"\b" + re.escape(p), L):
            cnt[fn][p] += 1
print(f"SASS summary of {os.path.relpath(lib, ROOT)} (sm_100a): cuobjdump -sass, mnemonic counts per kernel")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,")
print("HMMA = legacy mma.sync (TF32 in this library), LDGSTS = cp.async, UBLKCP = cp.async.bulk, RED/ATOM = global atomics\n")
for f, c in sorted(cnt.items(), key=lambda kv: -kv[1]["instructions"]):
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{name}: {c['instructions']} instructions ({c['instructions'] * 16 // 1024} KB)")
    print("    " + "  ".join(f"{p} {c[p]}" for p in pats if c[p]))
