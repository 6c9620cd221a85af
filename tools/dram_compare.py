"""Round-2 queue, step 3: DRAM bytes and time per trunk conv of the layer-by-layer path against the fused conv4+conv5 launch.

    python tools/dram_compare.py gpurun_out/r2_dram_fuse0.csv gpurun_out/r2_dram_fuse4.csv [windows side blocks]

Both CSVs are `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch lists of
tools/try_fused_ncu.py (25 windows of 276 x 276, 2 blocks).  Prints bytes per window pixel: the layer-by-layer rdb.conv4 +
rdb.conv5 pair against one rdb_fused_kernel launch — the 384 B/px dense-buffer read of conv5 should be gone if the skew keeps
it in L2."""
import collections
import csv
import sys

SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "second": 1e9}
T, RD, WR = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    byid = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = byid.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SC.get(r["Metric Unit"], 1)
    return list(byid.values())


def main():
    base, fused = launches(sys.argv[1]), launches(sys.argv[2])
    windows, side, blocks = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (25, 276, 2)
    px, n_rdb = windows * side * side, 3 * blocks
    tc = [k for k in base if "conv3x3_tc" in k["name"]][:n_rdb * 5]
    print(f"{windows} windows of {side}x{side}, {blocks} blocks; per RDB and window pixel (ncu: serialised, cold cache)")
    tot = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for i, k in enumerate(tc):
        a = tot[i % 5]
        a[0] += k[T]; a[1] += k[RD]; a[2] += k[WR]
    for j in range(5):
        t, r, w = tot[j]
        print(f"  layer-by-layer rdb.conv{j + 1}: {t / n_rdb / 1e3:8.1f} us  read {r / n_rdb / px:6.1f}  written {w / n_rdb / px:6.1f} B/px")
    p45 = [tot[3][i] + tot[4][i] for i in range(3)]
    fk = [k for k in fused if "rdb_fused" in k["name"]]
    if not fk:
        print("  no rdb_fused_kernel launch in", sys.argv[2])
        return 1
    ft, fr, fw = (sum(k[m] for k in fk) / len(fk) for m in (T, RD, WR))
    print(f"  conv4 + conv5, two launches : {p45[0] / n_rdb / 1e3:8.1f} us  read {p45[1] / n_rdb / px:6.1f}  written {p45[2] / n_rdb / px:6.1f} B/px")
    print(f"  rdb_fused_kernel x{len(fk):<3d}       : {ft / 1e3:8.1f} us  read {fr / px:6.1f}  written {fw / px:6.1f} B/px")
    print(f"  -> time x{ft / (p45[0] / n_rdb):.3f}, DRAM bytes x{(fr + fw) / ((p45[1] + p45[2]) / n_rdb):.3f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
