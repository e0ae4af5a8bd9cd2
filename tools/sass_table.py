"""SASS evidence: per-kernel counts of the Blackwell-native mnemonics in the in-tree library (B200_PROFILING.md, "What proves a
Blackwell-native kernel").   python tools/sass_table.py > profiles/r02_sass_table.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pnnp_b200", "libpnnp_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"),
            ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("FFMA2", r"\bFFMA2"), ("FADD2", r"\bFADD2"), ("DFMA", r"\bDFMA"),
            ("MUFU", r"\bMUFU"), ("IMAD.WIDE", r"\bIMAD\.WIDE"), ("LDG.128", r"\bLDG\.E\S*\.128"), ("STG.128", r"\bSTG\.E\S*\.128")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, total = collections.OrderedDict(), None, collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("pnnp::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            continue
        counts[cur]["instr"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    cols = ["instr"] + [n for n, _ in PATTERNS]
    print(f"# cuobjdump -sass pnnp_b200/libpnnp_b200.so (sm_100a), {len(counts)} kernels; counts of SASS instructions per kernel")
    print(f"# tcgen05.mma = UTCHMMA, TMA load = UTMALDG (store = UTMASTG), tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, mbarrier = SYNCS,")
    print(f"# legacy tensor path = HMMA (must be 0), packed fp32 pairs = FFMA2 / FADD2")
    print(f"{'kernel':64s} " + " ".join(f"{c:>9s}" for c in cols))
    groups = collections.OrderedDict()
    for k, c in counts.items():                     # fold the conv / wgrad template instantiations into one row per kernel family
        fam = re.sub(r"<.*", "<...>", k) if k.startswith(("conv_gemm_tc_kernel", "wgrad_nhwc_kernel")) else k
        g = groups.setdefault(fam, [0, collections.Counter()])
        g[0] += 1
        g[1].update(c)
        total.update(c)
    for fam, (n, c) in groups.items():
        label = fam if n == 1 else f"{fam} x{n} instantiations"
        print(f"{label[:64]:64s} " + " ".join(f"{c.get(col, 0):9d}" for col in cols))
    print(f"{'TOTAL':64s} " + " ".join(f"{total.get(col, 0):9d}" for col in cols))
    assert total["HMMA"] == 0, "legacy mma.sync code found"


if __name__ == "__main__":
    sys.exit(main())
