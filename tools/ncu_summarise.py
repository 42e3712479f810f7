"""Summarise .ncu-rep captures: key raw metrics per launch, the most-stalled SASS instructions (needs -lineinfo), and
-- with --traffic -- the per-kernel DRAM bytes per launch that bench.py reports as roofline.traffic.

    python tools/ncu_summarise.py gpurun_out/r02_prof_attn.ncu-rep [topn]          > profiles/r02_ncu_full_attn_summary.txt
    python tools/ncu_summarise.py --traffic profiles/traffic.json name=rep [...]   (name = kernel family of bench.py)
"""
import csv, io, json, subprocess, sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_xu.sum']


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(out)) if r]
    hi = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hi is None:
        return [], []
    hdr, units = rows[hi], rows[hi + 1] if hi + 1 < len(rows) else []
    body = rows[hi + 2:] if units and not units[0].isdigit() else rows[hi + 1:]
    return hdr, [(r, units) for r in body if len(r) == len(hdr)]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def summarise(rep, topn):
    hdr, body = raw_rows(rep)
    if not hdr:
        print(f"{rep}: no kernels captured")
        return
    for r, units in body:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:70s} {r[i]} {units[i] if units and i < len(units) and w != 'Kernel Name' else ''}")
        print()
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(src)) if r]
    hi = next((i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r), None)
    if hi is None:
        print("(no source page)")
        return
    hdr = rows[hi]
    ia, iex = hdr.index('Source'), hdr.index('Instructions Executed')
    isamp = hdr.index('Warp Stall Sampling (All Samples)') if 'Warp Stall Sampling (All Samples)' in hdr else None
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= max(ia, iex) or not r[iex].replace(",", "").isdigit():
            continue
        s = int(r[isamp].replace(",", "") or 0) if isamp is not None and len(r) > isamp and r[isamp].replace(",", "").isdigit() else 0
        data.append((s, r[ia].strip(), int(r[iex].replace(",", ""))))
    tot = sum(d[0] for d in data)
    print('total samples', tot, 'sass instructions', len(data), 'executed warp-instr', sum(d[2] for d in data))
    for i in sorted(range(len(data)), key=lambda i: -data[i][0])[:topn]:
        print(f"{i:5d} {data[i][0]:7d} {100 * data[i][0] / max(tot, 1):5.1f}%  ex={data[i][2]:9d}  {data[i][1][:100]}")


def traffic(out_path, pairs):
    res = {}
    for pair in pairs:
        name, rep = pair.split("=", 1)
        hdr, body = raw_rows(rep)
        if not hdr:
            continue
        ir, iw, ik, it = (hdr.index(k) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'Kernel Name', 'gpu__time_duration.sum'))
        caps = []
        for r, units in body:
            caps.append({"kernel": r[ik][:90], "dram_bytes": to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]),
                         "gpu_time_us": float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[it], 1)})
        if caps:
            e = res.setdefault(name, {"captures": [], "source": []})
            e["captures"] += caps
            e["source"].append(rep.split("/")[-1])
    for name, e in res.items():
        e["bytes_per_launch"] = sum(c["dram_bytes"] for c in e["captures"]) / len(e["captures"])
        e["source"] = ("ncu --set full --clock-control none (dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                       "cold L2, eager): " + ", ".join(e["source"]))
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps({k: round(v["bytes_per_launch"] / 1e6, 2) for k, v in res.items()}), "MB per launch ->", out_path)


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3:])
    else:
        summarise(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
