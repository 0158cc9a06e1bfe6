"""Counts of the instructions that prove which hardware path a kernel takes, per kernel of libmsvit.so
(cuobjdump -sass):  python tools/sass_counts.py > profiles/r2_sass_counts.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multi-state-vit_b200", "csrc", "libmsvit.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = {"UTCHMMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA", "UTCBAR (tcgen05.commit)": r"\bUTCBAR", "LDTM (tcgen05.ld)": r"\bLDTM",
        "STTM (tcgen05.st)": r"\bSTTM", "UTCATOMSWS (tmem alloc)": r"\bUTCATOMSWS", "UTMALDG (TMA tensor load)": r"\bUTMALDG",
        "UBLKCP (bulk copy)": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS", "HMMA (mma.sync)": r"\bHMMA", "CREDUX/REDUX": r"\bC?REDUX",
        "MATCH": r"\bMATCH", "MUFU.EX2": r"MUFU\.EX2", "LDL": r"\bLDL", "STL": r"\bSTL"}
counts = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
def dem(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
rows = sorted((re.sub(r"\(.*", "", dem(f)).replace("msvit::", "").replace("void ", ""), c) for f, c in counts.items())
keys = list(pats)
print("# SASS instruction counts per kernel (`cuobjdump -sass multi-state-vit_b200/csrc/libmsvit.so`, static counts)\n")
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---:|" * len(keys))
for n, c in rows:
    if sum(c.values()):
        print(f"| `{n}` | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
