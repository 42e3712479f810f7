"""Summarise an .ncu-rep: key raw metrics and the most-stalled SASS instructions (needs -lineinfo)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    for w in want:
        if w in hdr: print(f"  {w:70s} {r[hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isamp, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
data = [(int(r[isamp] or 0), r[ia].strip(), int(r[iex] or 0)) for r in rows[2:] if len(r) > isamp]
tot = sum(d[0] for d in data)
print('total samples', tot, 'sass instructions', len(data), 'executed warp-instr', sum(d[2] for d in data))
for i in sorted(range(len(data)), key=lambda i: -data[i][0])[:topn]:
    print(f"{i:5d} {data[i][0]:7d} {100*data[i][0]/max(tot,1):5.1f}%  ex={data[i][2]:9d}  {data[i][1][:100]}")
