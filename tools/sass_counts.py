#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tensor-core / TMEM / async-copy paths.

usage: python tools/sass_counts.py > profiles/rNN_sass_counts.md   (runs cuobjdump on the in-tree .so)
"""
import collections
import re
import subprocess
from pathlib import Path

SO = Path(__file__).resolve().parent.parent / "lkpy_b200" / "csrc" / "liblkpy_b200.so"
PATS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "LDGSTS", "SYNCS", "ATOMS", "ATOMG", "REDUX", "FFMA2",
        "FFMA", "MUFU.RSQ", "BAR.SYNC", "ELECT"]  # fmt: skip


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True, check=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            counts[cur]["_total"] += 1
            for p in PATS:
                if m.group(1).startswith(p):
                    counts[cur][p] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence — `cuobjdump -sass lkpy_b200/csrc/liblkpy_b200.so` (sm_100a), instruction counts per kernel\n")
    print("`UTCHMMA` = tcgen05.mma (kind::f16 and kind::tf32), `UTCBAR` = tcgen05.commit, `LDTM` / `STTM` = tcgen05.ld / st,")
    print("`LDGSTS` = cp.async, `UBLKCP` = cp.async.bulk (the TMA engine's 1-D bulk copy), `SYNCS` = mbarrier operations.\n")
    print("| kernel | SASS instr | " + " | ".join(PATS) + " |")
    print("|---|---:|" + "---:|" * len(PATS))
    for key, name in sorted(zip(counts, names), key=lambda kn: -counts[kn[0]]["_total"]):
        c = counts[key]
        if c["_total"] < 50:
            continue
        short = re.sub(r"\(.*$", "", name).replace("void ", "")
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[p]) if c[p] else "" for p in PATS) + " |")


if __name__ == "__main__":
    main()
