"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu_source_lines.py dump.csv [launch_index ...]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = [int(a) for a in sys.argv[2:]] or [0]
# a dump holds, per profiled launch, one block per source file
launch, seen_files, cur_file, hdr, blocks = -1, set(), None, None, collections.defaultdict(list)
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        if cur_file in seen_files or launch < 0:
            launch += 1; seen_files = set()
        seen_files.add(cur_file); hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":          # per-source-line summary rows
        blocks[launch].append((cur_file, r, hdr))
for li in want:
    data = blocks[li]
    if not data: print("no data for launch", li); continue
    hi = data[0][2]; ii = hi.index("Instructions Executed"); si = hi.index("# Samples")
    tot = sum(float(r[ii] or 0) for _, r, _ in data); tots = sum(float(r[si] or 0) for _, r, _ in data)
    print(f"== launch {li}: {tot/1e6:.1f} M warp-instructions, {tots:.0f} samples")
    for f, r, _ in sorted(data, key=lambda x: -float(x[1][ii] or 0))[:26]:
        print(f"{float(r[ii])/tot*100:5.1f}% inst {float(r[si] or 0)/tots*100:5.1f}% smp  {f}:{r[0]:>4s}  {r[1].strip()[:100]}")
