"""usage: ncu -i rep.ncu-rep --page source --csv --print-source sass | python tools/sass_opcode_mix.py [items]
Executed warp instructions by opcode (and by opcode x predicate-free mnemonic stem) from ncu's SASS source page."""
import collections
import csv
import sys

items = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Instructions Executed" in c for c in r):
        hdr, start = r, i + 1
        break
if hdr is None:
    sys.exit("no source table found")
isrc = hdr.index("Source")
iex = [i for i, c in enumerate(hdr) if c.strip() == "# Instructions Executed" or c.strip() == "Instructions Executed"][0]
ist = [i for i, c in enumerate(hdr) if "Warp Stall Sampling (All" in c]
agg = collections.Counter()
stall = collections.Counter()
tot = 0
lines = []
for r in rows[start:]:
    if len(r) != len(hdr):
        continue
    try:
        n = int(float(r[iex].replace(",", "")))
    except ValueError:
        continue
    txt = r[isrc].strip()
    toks = txt.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    stem = op.split(".")[0]
    agg[stem] += n
    tot += n
    if ist:
        try:
            stall[stem] += int(float(r[ist[0]].replace(",", "")))
        except ValueError:
            pass
    lines.append((n, txt))
print(f"{tot} executed warp instructions = {tot / items:.0f} per item")
st = sum(stall.values()) or 1
for op, n in agg.most_common(40):
    print(f"  {op:12s} {n:12d} {100.0 * n / tot:5.1f} %   {n / items:8.1f} per item   stall samples {100.0 * stall[op] / st:5.1f} %")
print("hottest single instructions:")
for n, txt in sorted(lines, reverse=True)[:25]:
    print(f"  {n:10d}  {txt[:110]}")
