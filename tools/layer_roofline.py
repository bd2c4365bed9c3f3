"""Per-launch roofline of one ResNet-50 step from an ncu launch list (gpu__time_duration.sum csv).

usage: python tools/layer_roofline.py gpurun_out/<tag>_launches.csv [batch]
Every launch gets max(FLOPs / sustained bf16 peak, algorithmic bytes / measured HBM copy bandwidth)
from MEASURED_PEAKS.json next to its measured (cold-cache, serialised) duration.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
TF, GB = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
names = [r[4] for r in rows]
t = [float(r[-1]) / 1e3 for r in rows]
starts = [i for i, n in enumerate(names) if "pack_stem" in n]
s = starts[1]
meas = t[s:s + 57]
L = []


def conv(name, hin, cin, cout, k, stride, res=False):
    ho = hin // stride
    M = B * ho * ho
    inb = B * hin * hin * cin * 2 if not (k == 1 and stride == 2) else B * ho * ho * cin * 2
    L.append((name, M, k * k * cin, cout, inb, M * cout * 2, M * cout * 2 if res else 0))
    return ho


L.append(("pack", 0, 0, 0, B * 3 * 224 * 224 * 4, B * 230 * 232 * 8 * 2, 0))
L.append(("stem", B * 112 * 112, 147, 64, B * 230 * 232 * 8 * 2, B * 112 * 112 * 64 * 2, 0))
L.append(("maxpool", 0, 0, 0, B * 112 * 112 * 64 * 2, B * 56 * 56 * 64 * 2, 0))
h, cin = 56, 64
for li, (w, n, st) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]):
    for b in range(n):
        s_ = st if b == 0 else 1
        if b == 0:
            conv(f"l{li+1}.{b}.down", h, cin, w * 4, 1, s_)
        conv(f"l{li+1}.{b}.c1", h, cin, w, 1, 1)
        h2 = conv(f"l{li+1}.{b}.c2", h, w, w, 3, s_)
        conv(f"l{li+1}.{b}.c3", h2, w, w * 4, 1, 1, res=True)
        h, cin = h2, w * 4
L.append(("avgpool", 0, 0, 0, B * 49 * 2048 * 2, B * 2048 * 2, 0))
L.append(("fc", B, 2048, 1000, B * 2048 * 2, B * 1000 * 4, 0))
tot_m = tot_r = 0
print(f"{'launch':12s} {'kernel':22s} {'M':>7s} {'K':>5s} {'N':>5s} {'meas us':>8s} {'tensor':>7s} {'hbm':>7s} {'eff%':>6s} {'loss us':>8s}")
for (name, M, K, N, ib, ob, rb), m, kn in zip(L, meas, names[s:s + 57]):
    fl = 2 * M * K * N
    t_t, t_h = fl / TF * 1e6, (ib + ob + rb) / GB * 1e6
    r = max(t_t, t_h)
    tot_m += m
    tot_r += r
    kn = kn.split("(")[0].replace("void ", "").replace("eqxv::", "")
    print(f"{name:12s} {kn:22s} {M:7d} {K:5d} {N:5d} {m:8.1f} {t_t:7.1f} {t_h:7.1f} {r/m*100:6.1f} {m-r:8.1f}")
print(f"step: measured {tot_m:.1f} us, layer-wise roofline {tot_r:.1f} us ({tot_r/tot_m*100:.1f}%)")
