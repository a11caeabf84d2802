"""Blackwell evidence for profiles/: per kernel of the built library, how many tcgen05 / TMEM / TMA instructions its SASS holds
(cuobjdump -sass; the PTX names never appear in SASS: tcgen05.mma -> UTC*MMA, tcgen05.ld / st -> LDTM / STTM,
cp.async.bulk.tensor -> UTMALDG / UTMASTG, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, mma.sync -> HMMA).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mcm_b200", "_C", "libmcm_b200.so")
PATS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.TANH", "LDGSTS", "REDUX"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            counts[cur]["_instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(2)
        counts[cur]["_instructions"] += 1
        for p in PATS:
            if p == "UTCHMMA":
                if op.startswith("UTCHMMA") and ".2CTA" not in op:
                    counts[cur][p] += 1
            elif op.startswith(p) or (p == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                counts[cur][p] += 1
    names = list(counts)
    try:
        dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dem))
    except OSError:
        pass
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)")
    print("# columns: " + " ".join(PATS) + " | instructions | kernel")
    tot = collections.Counter()
    for n, c in counts.items():
        short = re.sub(r"\(.*", "", demangle.get(n, n))
        short = short.replace("void mcm::", "")
        print(" ".join(f"{c[p]:5d}" for p in PATS) + f" | {c['_instructions']:6d} | {short}")
        tot.update(c)
    print(" ".join(f"{tot[p]:5d}" for p in PATS) + f" | {tot['_instructions']:6d} | TOTAL ({len(counts)} kernels)")


if __name__ == "__main__":
    sys.exit(main())
