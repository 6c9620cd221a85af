"""Per-layer summary of an ncu launch list of one bench step (gpu__time_duration.sum [+ dram bytes] csv).
usage: python tools/launch_report.py <csv> <windows> <window_side> [traffic_json_out]"""
import collections
import csv
import json
import sys

path, windows, side = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
byid = collections.OrderedDict()
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "second": 1e9}
for r in csv.DictReader(lines):
    d = byid.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SC.get(r["Metric Unit"], 1)
ks = list(byid.values())
i0 = [i for i, k in enumerate(ks) if "conv_first" in k["name"]][0]
step = ks[i0:i0 + 354]
tc = [k for k in step if "conv3x3_tc" in k["name"] or "conv3x3_roll" in k["name"]]
T, RD, WR = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
has_dram = RD in tc[0]
px = windows * side * side
fl = [36864, 55296, 73728, 92160, 221184]
print(f"one step of bench.py on {windows} windows of {side}x{side} (ncu: serialised launches, cold cache — compare shares)")
print(f"conv_first_kernel {step[0][T]/1e3:.1f} us")
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0, 1e30, 0.0])
for i, k in enumerate(tc[:345]):
    a = agg[i % 5]
    a[0] += k[T]; a[3] += 1; a[4] = min(a[4], k[T]); a[5] = max(a[5], k[T])
    if has_dram:
        a[1] += k[RD]; a[2] += k[WR]
tot = sum(k[T] for k in tc)
for j in range(5):
    t, r, w, n, mn, mx = agg[j]
    extra = f"  dram {r/n/px:6.1f} B/px read {w/n/px:6.1f} B/px written" if has_dram else ""
    print(f"rdb.conv{j+1} x{n}: {t/n/1e3:7.1f} us avg (min {mn/1e3:.1f} max {mx/1e3:.1f})  {px*fl[j]/(t/n)/1e3:5.0f} TFLOP/s algorithmic  {100*t/tot:4.1f}% of conv time{extra}")
fls = [73728, 294912, 1179648, 1179648, 55296]
for nm, k, f in zip(["conv_body", "conv_up1", "conv_up2", "conv_hr", "conv_last"], tc[345:], fls):
    extra = f"  dram {k[RD]/px:7.1f} read {k[WR]/px:7.1f} written B/LR-px" if has_dram else ""
    print(f"{nm}: {k[T]/1e3:8.1f} us  {px*f/k[T]/1e3:5.0f} TFLOP/s algorithmic  {100*k[T]/tot:4.1f}% of conv time{extra}")
print(f"all {len(tc)} tensor-core conv launches: {tot/1e6:.2f} ms -> {px*35853696/tot/1e3:.0f} TFLOP/s algorithmic")
if has_dram:
    b = sum(k[RD] + k[WR] for k in tc)
    print(f"DRAM traffic of those launches: {b/1e9:.1f} GB ({b/tot/1e3:.2f} TB/s while running, {px*35853696/b:.0f} FLOP/B)")
    if len(sys.argv) > 4:
        json.dump({"tc_launches": len(tc), "time_ms": tot / 1e6, "dram_bytes": b, "windows": windows, "side": side}, open(sys.argv[4], "w"))
for k in step:
    if "conv" not in k["name"]:
        print(f"{k['name'][:70]} grid {k['grid']}: {k[T]/1e3:.1f} us")
