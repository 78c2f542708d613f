"""Attributes an ncu report's per-SASS-instruction counts and stall samples to CUDA source lines.
usage: ncu_lines.py report.ncu-rep cubin mangled_kernel_name [topN]"""
import collections, csv, io, re, subprocess, sys
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\." + re.escape(kern), l))
addr2line, cur = {}, None
for l in dis[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ai, ii, si = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
inst, smp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        a, n, s = int(r[ai], 16), int(r[ii]), int(r[si])
    except Exception:
        continue
    if base is None:
        base = a
    k = addr2line.get(a - base, ("?", 0))
    inst[k] += n
    smp[k] += s
ti, ts = sum(inst.values()), sum(smp.values())
print(f"total warp instructions {ti}, samples {ts}")
for k, s in smp.most_common(top):
    print(f"{s / ts * 100:5.1f}% samples  {inst[k] / ti * 100:5.1f}% inst  {k[0]}:{k[1]}")
