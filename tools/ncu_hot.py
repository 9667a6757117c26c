#!/usr/bin/env python
"""Top stall-sample instructions of each kernel in an .ncu-rep (source page, SASS view)."""
import csv, io, subprocess, sys
src, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
if src.endswith(".gz"):      # `ncu -i <rep> --page source --csv | gzip` exported on the GPU box
    import gzip
    out = gzip.open(src, "rt").read()
else:
    out = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for bi, b in enumerate(blocks):
    h = b["hdr"]
    ia, isrc, iall, iex = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    tot = sum(int(r[iall] or 0) for r in b["rows"])
    print(f"== kernel {bi}: {b['name'][:80]}  total samples {tot}, {len(b['rows'])} instrs")
    rows = sorted(enumerate(b["rows"]), key=lambda kv: -int(kv[1][iall] or 0))[:topn]
    for idx, r in sorted(rows):
        print(f"  [{idx:5d}] {100*int(r[iall] or 0)/max(tot,1):5.1f}%  exec={r[iex]:>9}  {r[isrc].strip()[:90]}")
