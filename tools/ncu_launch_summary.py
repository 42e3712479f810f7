"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv):
    python tools/ncu_launch_summary.py gpurun_out/r02w_ncu_launches_XL8.csv > profiles/r02w_ncu_launches_XL8_summary.txt"""
import csv, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')) if r]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    name = re.sub(r"^(void )?(ma3::)?", "", r[ik])
    name = re.split(r"[<(]", name)[0]
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}[r[iu]]
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += us
tot = sum(v[1] for v in agg.values())
print(f"# one eager step of bench.py (XL, 8 prompts): {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms serialised (ncu, cold caches)")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:36s} {n:5d} launches {us / 1e3:9.3f} ms {100 * us / tot:6.2f} % {us / n:9.2f} us/launch")
