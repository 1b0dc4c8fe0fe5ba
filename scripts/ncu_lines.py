"""Attribute ncu warp-stall samples / executed instructions of one kernel to CUDA source lines.
usage: python scripts/ncu_lines.py <report.ncu-rep> <cubin from `cuobjdump -xelf all lib.so`> <kernel-name-substring> <source-file>"""
import collections, csv, io, re, subprocess, sys
rep, cubin, kname, srcfile = sys.argv[1:5]
THR = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the .text section of the kernel: remember the last "//## File ..., line N" (innermost non-inlined: take first of a run)
line_of = {}; cur = None; inside = False
for l in dis:
    if l.startswith("//") and ".text." in l: inside = kname in l
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # prefer the location inside our source file (inlined-at chains list the callee first)
        if srcfile in m.group(1): cur = int(m.group(2))
        else:
            m2 = re.search(r'inlined at "[^"]*' + re.escape(srcfile) + r'", line (\d+)', l)
            if m2: cur = int(m2.group(1))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and cur is not None: line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ia, iss, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for x in rows[2:]:
    if len(x) > 10 and x[0].startswith("0x"): data.append((int(x[ia], 16), int(x[iss]), int(x[iex])))
    elif data: break
base = data[0][0]
agg = collections.defaultdict(lambda: [0, 0])
for a, sm, ex in data:
    ln = line_of.get(a - base, -1); agg[ln][0] += sm; agg[ln][1] += ex
ts = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
text = open(srcfile).read().splitlines()
print(f"samples {ts}, executed {te}; lines with >=1% of either:")
for ln in sorted(agg):
    sm, ex = agg[ln]
    if sm >= ts * THR or ex >= te * THR:
        print(f"  L{ln:4d} smp {sm/ts*100:5.1f}%  exec {ex/te*100:5.1f}%  | {text[ln-1].strip()[:90] if 0 < ln <= len(text) else '?'}")
