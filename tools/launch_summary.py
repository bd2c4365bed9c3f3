"""Summarise an ncu per-launch metric csv (tools/gpu_round.sh): one step of the model, per kernel family:
launches, total time, share of the step, DRAM bytes and achieved DRAM GB/s, tensor-pipe activity.

usage: python tools/launch_summary.py <csv> <first-kernel-substring> [--list]
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
first = sys.argv[2]
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(int(r[0]), {"name": r[4], "grid": r[8]})
    d[r[12]] = float(r[14].replace(",", "")) if r[14] not in ("", "n/a") else 0.0
    d["unit:" + r[12]] = r[13]
L = list(by_id.values())
starts = [i for i, d in enumerate(L) if first in d["name"]]
s, e = starts[1], starts[2] if len(starts) > 2 else len(L)
step = L[s:e]


def t_us(d):
    v, u = d.get("gpu__time_duration.sum", 0.0), d.get("unit:gpu__time_duration.sum", "ns")
    return v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)


def mb(d, k):
    v, u = d.get(k, 0.0), d.get("unit:" + k, "byte")
    return {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(u, v)


agg = collections.OrderedDict()
for d in step:
    k = d["name"].split("(")[0].replace("void ", "").replace("eqxv::", "")
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += t_us(d)
    a[2] += mb(d, "dram__bytes_read.sum")
    a[3] += mb(d, "dram__bytes_write.sum")
    a[4] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t_us(d)
tot = sum(a[1] for a in agg.values())
print(f"one step: {len(step)} launches, {tot:.1f} us (ncu: serialised, cold cache)")
print(f"{'kernel':44s} {'n':>4s} {'us':>9s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s} {'tensor%':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} {a[0]:4d} {a[1]:9.1f} {a[1]/tot*100:5.1f}% {a[2]:9.1f} {a[3]:9.1f} {(a[2]+a[3])/a[1]*1e3:7.0f} {a[4]/a[1]:7.1f}")
print(f"{'total':44s} {len(step):4d} {tot:9.1f}        {sum(a[2] for a in agg.values()):9.1f} {sum(a[3] for a in agg.values()):9.1f}")
if "--list" in sys.argv:
    for i, d in enumerate(step):
        k = d["name"].split("(")[0].replace("void ", "").replace("eqxv::", "")
        print(f"{i:3d} {k:40s} {d['grid']:16s} {t_us(d):8.1f} us  rd {mb(d,'dram__bytes_read.sum'):8.1f} wr {mb(d,'dram__bytes_write.sum'):8.1f} MB "
              f"{(mb(d,'dram__bytes_read.sum')+mb(d,'dram__bytes_write.sum'))/t_us(d)*1e3:6.0f} GB/s tensor {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',0):5.1f}%")
