"""Summarises an ncu launch list (gpu__time_duration.sum csv) of one bench step per layer type."""
import collections
import csv
import sys

path = sys.argv[1]
windows, side = int(sys.argv[2]), int(sys.argv[3])
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)]
first = [i for i, r in enumerate(rows) if "conv_first" in r[0]]
i0 = first[0]
tc = [r for r in rows[i0 + 1:] if "conv3x3_tc" in r[0]][:350]
px = windows * side * side
fl = [36864, 55296, 73728, 92160, 221184]
agg = collections.defaultdict(list)
for i, r in enumerate(tc[:345]):
    agg[i % 5].append(r[2])
tot = 0
print(f"conv_first {rows[i0][2]/1e3:.1f} us")
for k in range(5):
    t = sum(agg[k]) / len(agg[k])
    tot += sum(agg[k])
    print(f"rdb.conv{k+1}: {t/1e3:8.1f} us avg (min {min(agg[k])/1e3:.1f} max {max(agg[k])/1e3:.1f}) -> {px*fl[k]/t/1e3:6.0f} TFLOP/s algorithmic, {sum(agg[k])/1e6:6.2f} ms total")
fls = [73728, 294912, 1179648, 1179648, 55296]
for nm, r, f in zip(["conv_body", "conv_up1", "conv_up2", "conv_hr", "conv_last"], tc[345:], fls):
    tot += r[2]
    print(f"{nm}: {r[2]/1e3:8.1f} us -> {px*f/r[2]/1e3:6.0f} TFLOP/s algorithmic")
print(f"all tensor-core conv launches: {tot/1e6:.2f} ms -> {px*35853696/tot/1e3:.0f} TFLOP/s algorithmic (ncu-serialised, cold cache)")
for r in rows[i0:]:
    if "conv" not in r[0]:
        print(f"{r[0][:60]} {r[1]} {r[2]/1e3:.1f} us")
