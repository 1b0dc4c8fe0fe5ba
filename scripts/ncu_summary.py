"""Summarise an .ncu-rep (raw + source pages) for one kernel: key metrics, stall mix, hot SASS regions.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [bucket_bytes]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; bucket = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x400
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, r = rows[0], rows[1], rows[2]
print("kernel:", r[hdr.index("Kernel Name")][:100])
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in hdr: print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: float(r[i]) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and r[i]}
tot = sum(st.values()) or 1
print("  stall samples:", ", ".join(f"{k} {v/tot*100:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ia, isrc, iss, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for x in rows[2:]:
    if len(x) > 10 and x[0].startswith("0x"): data.append((int(x[ia], 16), x[isrc].strip(), int(x[iss]), int(x[iex])))
    elif data: break
base = data[0][0]; ts = sum(d[2] for d in data); te = sum(d[3] for d in data)
print(f"  SASS instrs {len(data)}, samples {ts}, warp-instructions executed {te}")
b = collections.OrderedDict()
for a, s, sm, ex in data:
    v = b.setdefault((a - base) // bucket, [0, 0]); v[0] += sm; v[1] += ex
for k, (sm, ex) in b.items():
    if sm > ts * 0.015 or ex > te * 0.015: print(f"   +{k*bucket:#07x}: samples {sm/ts*100:5.1f}%  executed {ex/te*100:5.1f}%")
print("  hottest instructions:")
for a, s, sm, ex in sorted(data, key=lambda d: -d[2])[:14]: print(f"   +{a-base:#07x} {sm/ts*100:5.1f}% ex={ex:10d}  {s[:80]}")
