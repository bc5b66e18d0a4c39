"""Per-kernel table of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), UTCHMMA.2CTA (cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store),
UTCBAR (tcgen05.commit), STSM / LDSM (stmatrix / ldmatrix), SYNCS (mbarrier).    python tools/sass_table.py > profiles/r02_sass_table.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audioset-convnext-inf_b200", "libacx.so")
PATS = [("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTCBAR", r"\bUTCBAR"), ("STSM", r"\bSTSM"),
        ("SYNCS", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"), ("MUFU.TANH", r"MUFU\.TANH")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATS:
            if re.search(pat, line):
                counts[cur][name] += 1
    stamp = open(os.path.join(ROOT, "audioset-convnext-inf_b200", ".libacx.stamp")).read().strip()[:16]
    print(f"# SASS mnemonic counts per kernel of libacx.so (source digest {stamp}…)\n")
    print("`cuobjdump -sass audioset-convnext-inf_b200/libacx.so`, one row per kernel that has any of them.\n")
    print("| kernel | " + " | ".join(n for n, _ in PATS) + " |")
    print("|---|" + "---|" * len(PATS))
    tot = collections.Counter()
    for k, c in counts.items():
        if not any(c[n] for n, _ in PATS[:9]):
            continue
        tot.update(c)
        print(f"| `{k}` | " + " | ".join(str(c[n]) if c[n] else "" for n, _ in PATS) + " |")
    print("| **total** | " + " | ".join(str(tot[n]) for n, _ in PATS) + " |")


if __name__ == "__main__":
    sys.exit(main())
