"""Per-kernel SASS mnemonic table of the shipped liboai_b200.so (cuobjdump -sass), the evidence file
profiles/rNN_sass_mnemonics.txt:  python scripts/sass_table.py > profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oai_analysis_2_b200", "liboai_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "SYNCS", "HMMA", "STG", "LDG", "LDS", "STS", "DFMA", "FFMA",
        "ATOMG", "RED", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True,
                              text=True).stdout.splitlines()
    names = iter(demangle)
    table, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            full = next(names)
            short = full.replace("(anonymous namespace)::", "").replace("oai::", "")
            short = re.sub(r"^void ", "", re.sub(r"\(.*$", "", short))
            cur = table.setdefault(short, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["total"] += 1
    print("# Per-kernel SASS mnemonic counts of the shipped liboai_b200.so (cuobjdump -sass, sm_100a); scripts/sass_table.py.")
    print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA tiled load), UBLKCP = "
          "cp.async.bulk,")
    print("# LDTM = tcgen05.ld, SYNCS = mbarrier ops, HMMA = mma.sync (legacy warp-level tensor path).")
    print("%-58s" % "kernel" + "".join("%9s" % c for c in COLS) + "    total")
    for k, c in table.items():
        print("%-58s" % k[:58] + "".join("%9d" % c[x] for x in COLS) + "%9d" % c["total"])


if __name__ == "__main__":
    sys.exit(main())
