#!/usr/bin/env python
"""Stall samples per CUDA source line: joins an ncu SASS-page CSV with nvdisasm -g line info of the same cubin.
usage: stalls_by_line.py X.ncu-rep libuaisa_b200.so 'uaisa_env_kernelILi4ELi128ELb1ELb0' [N]"""
import re, csv, collections, subprocess, sys, tempfile, os, io, glob
rep, lib, kname = sys.argv[1:4]; topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
seq = []
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
    if kname not in txt or seq:
        continue
    inside = False; cur = ("?", 0)
    for ln in txt.split("\n"):
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            inside = kname in ln; continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            seq.append((cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr = None; data = []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and r and len(r) == len(hdr):
        data.append(r)
if len(data) > len(seq) and len(data) % len(seq) == 0:      # several launches of the kernel in one report: take the last one
    data = data[-len(seq):]
assert len(seq) == len(data), (len(seq), len(data), "the report and the library are different builds")
ss = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by = collections.Counter(); byi = collections.Counter(); byst = collections.defaultdict(collections.Counter)
for (loc, _), r in zip(seq, data):
    by[loc] += int(r[ss] or 0); byi[loc] += int(r[ie] or 0)
    for c in cols:
        byst[loc][hdr[c][6:]] += int(r[c] or 0)
tot = sum(by.values()); toti = sum(byi.values())
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3drvo-marl-collisionavoidance_b200", "csrc")
src = {f: open(os.path.join(root, f)).read().split("\n") for f in os.listdir(root) if f.endswith((".cu", ".cuh"))}
print("kernel %s: %d samples, %d warp-instructions" % (kname, tot, toti))
for loc, c in by.most_common(topn):
    top = ", ".join("%s %d" % kv for kv in byst[loc].most_common(3))
    text = src[loc[0]][loc[1] - 1].strip()[:90] if loc[0] in src and loc[1] <= len(src[loc[0]]) else ""
    print("%-24s samples %4.1f%%  instr %4.1f%%  [%s]  %s" % ("%s:%d" % loc, 100.0 * c / tot, 100.0 * byi[loc] / toti, top, text))
