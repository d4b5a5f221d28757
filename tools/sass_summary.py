"""Opcode census of the built library per kernel (`cuobjdump -sass`), kept under profiles/ as the evidence that the hot
kernels are Blackwell-native: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA, HMMA = legacy
mma.sync (must be absent); the `.2CTA` forms (UTCHMMA.2CTA, UTMALDG.*.2CTA, UTCBAR.2CTA.MULTICAST) are the CTA-pair GEMM.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "videopainter_b200", "csrc", "libvp_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU", "FFMA2", "F2FP", "FMNMX3", "STS", "LDS",
         "HMMA", "HGMMA", "LDL", "STL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, census = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            tag = re.search(r"__N__[0-9a-f]+_\d+_([a-z0-9_]+?)_cu_", m.group(1))             # source file of an anonymous-namespace kernel
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "") + (f" [{tag.group(1)}.cu]" if tag else "")
            census[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            census[cur][m.group(1)] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (sm_100a); columns: " + " ".join(WATCH))
    tot = collections.Counter()
    for k, c in census.items():
        tot.update(c)
        print(f"{k[:70]:70s} instr={sum(c.values()):6d} " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
    print("# whole library: " + " ".join(f"{w}={tot[w]}" for w in WATCH))
    full = collections.Counter(re.findall(r"\b(UTCHMMA[.\w]*|UTMALDG[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*)", txt))
    print("# tcgen05 / TMA opcodes with modifiers: " + " ".join(f"{k}={v}" for k, v in sorted(full.items())))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core instructions found"


if __name__ == "__main__":
    main()
