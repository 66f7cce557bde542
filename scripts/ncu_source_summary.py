"""Summarise `ncu --page source --csv` output: opcode mix and the hottest SASS lines (run on the CPU box)."""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[0].startswith("0x")]
iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
toti = sum(int(r[iI]) for r in data)
print("total samples", tot, "warp-inst", toti, "sass lines", len(data))
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    s = r[isrc].strip().split()
    op = s[1] if s[0].startswith("@") else s[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST")) else op.split(".")[0]
    ops[op] += int(r[iI])
    samp[op] += int(r[iS])
for op, n in ops.most_common(28):
    print(f"{op:14s} {n:12d} {100 * n / toti:5.1f}%  samples {100 * samp[op] / tot:5.1f}%")
print()
for r in sorted(data, key=lambda r: -int(r[iS]))[:top]:
    stalls = {h[6:]: int(r[i]) for i, h in enumerate(hdr)
              if h.startswith("stall_") and "(Not" not in h and r[i] not in ("", "0")}
    print(r[iS], r[isrc].strip()[:64], stalls)
