#!/usr/bin/env python
"""Aggregate an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass) by CUDA source
line, using nvdisasm -g line info of the same cubin.  Usage: sass_by_line.py sass.csv kernels.sass 'ILi4ELi256E' """
import csv, re, sys, collections
sass_csv, disasm, kname = sys.argv[1:4]
# --- nvdisasm: ordered list of (file,line,opcode) for the kernel
lines = open(disasm).read().split("\n")
inside = False; cur = ("?", 0); seq = []
for ln in lines:
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip()); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        seq.append((cur, m.group(2).strip()))
# --- ncu rows for that kernel
rows = list(csv.reader(open(sass_csv)))
blocks = []; cur_rows = None; hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_rows = []; blocks.append((r[1], cur_rows)); continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur_rows is not None and r:
        cur_rows.append(r)
name, data = blocks[0]
ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples"); src = hdr.index("Source")
print("kernel:", name, "sass rows:", len(data), "disasm instrs:", len(seq))
n = min(len(data), len(seq))
by_line = collections.Counter(); by_line_s = collections.Counter(); by_op = collections.Counter(); tot = 0; tots = 0
for i in range(n):
    (loc, op) = seq[i]
    c = int(data[i][ie] or 0); s = int(data[i][ss] or 0)
    by_line[loc[:2]] += c; by_line_s[loc[:2]] += s; tot += c; tots += s
    by_op[data[i][src].split()[0] if not data[i][src].strip().startswith("@") else data[i][src].split()[1]] += c
print("total warp-instructions executed:", tot, "samples:", tots)
print("\n-- top source lines by executed warp-instructions")
for (loc, c) in by_line.most_common(45):
    print("%-22s %6d  %5.1f%%  samples %5.1f%%" % ("%s:%d" % loc, c // 1, 100.0 * c / tot, 100.0 * by_line_s[loc] / max(tots, 1)))
print("\n-- top opcodes")
for (op, c) in by_op.most_common(25):
    print("%-12s %5.1f%%" % (op, 100.0 * c / tot))
