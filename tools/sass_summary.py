"""Per-kernel SASS mnemonic counts of the shipped library (tcgen05 / TMEM / TMA evidence, and the absence of legacy HMMA).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ctrlora_b200", "lib", "libctrlora_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "MUFU.EX2", "SYNCS",
        "ACQBULK", "UCGABAR_ARV", "CCTL", "ATOMS", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(name.split("(")[0].replace("void ", ""), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["instructions"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k.endswith(".2CTA") and op.startswith("UTCHMMA") and ".2CTA" in op):
                    cur[k] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}  (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':64s} {'instr':>7s} " + " ".join(f"{k:>9s}" for k in KEYS))
    tot = collections.Counter()
    for name, c in per.items():
        print(f"{name[:64]:64s} {c['instructions']:7d} " + " ".join(f"{c[k]:9d}" for k in KEYS))
        tot.update(c)
    print(f"{'TOTAL':64s} {tot['instructions']:7d} " + " ".join(f"{tot[k]:9d}" for k in KEYS))


if __name__ == "__main__":
    main()
