"""Per-kernel SASS mnemonic counts of libvllm_b200.so (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma,
UTMALDG / UTMASTG / UBLKCP = TMA, LDTM / STTM = tcgen05.ld/st, HMMA = legacy mma.sync).  No GPU needed.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "visionllm_b200", "lib", "libvllm_b200.so")
PATS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDGSTS",
        "FHFMA", "LDG.E.128", "LDS.128", "SHFL", "RED.", "MULTIMEM"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", txt)
    if names:
        dm = subprocess.run(["cu++filt"] + sorted(set(names)), capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(sorted(set(names)), dm))
    cur, counts, total = None, collections.OrderedDict(), {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            total[cur] += 1
            for p in PATS:
                if p in line:
                    counts[cur][p] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} -- instruction counts per kernel (static SASS, sm_100a)")
    print("# " + " ".join(PATS))
    for k, c in counts.items():
        name = re.sub(r"\(.*", "", demangle.get(k, k)).replace("void ", "")
        hits = " ".join(f"{p}={c[p]}" for p in PATS if c[p])
        print(f"{name[:110]:110s} instr={total[k]:6d}  {hits}")


if __name__ == "__main__":
    main()
