"""Join ncu per-SASS-instruction counts (source page, sass view) with nvdisasm -g line info: executed warp
instructions per CUDA source line.  usage: sass_lines.py <report.ncu-rep> <kernel substring> <nvdisasm -g -c output> [top]"""
import collections, csv, io, re, subprocess, sys
rep, kname, sass = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ncu_name = sys.argv[5] if len(sys.argv) > 5 else kname
# line info per instruction index within the function
lines = open(sass).read().split("\n")
start = None
for i, l in enumerate(lines):
    if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"):
        start = i; break
cur = None; per_instr = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        f = m.group(1).split("/")[-1]; cur = (f, int(m.group(2)), (m.group(3) or "").split("/")[-1], int(m.group(4) or 0)); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l): per_instr.append((cur, l.strip()))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blk = [b for b in txt.split('"Kernel Name",')[1:] if ncu_name in b.split("\n")[0]][0]
rows = list(csv.reader(io.StringIO('"Kernel Name",' + blk))); h = rows[1]
ie = h.index("Instructions Executed"); isamp = h.index("# Samples")
cnt = [(int(r[ie]), int(r[isamp])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
assert abs(len(cnt) - len(per_instr)) < 8, (len(cnt), len(per_instr))
agg = collections.Counter(); sagg = collections.Counter(); tot = 0
for (li, _), (n, s) in zip(per_instr, cnt):
    key = li if li is None else ((li[2], li[3], "<-" + li[0] + ":" + str(li[1])) if li[2] else (li[0], li[1], ""))
    agg[key] += n; sagg[key] += s; tot += n
print("total warp instructions", tot)
for k, n in agg.most_common(top):
    print(f"{100*n/tot:5.1f}% {n:12d} samples {sagg[k]:7d}  {k}")
