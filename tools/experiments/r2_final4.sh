# Round 2, last GPU seconds: cfg3 (EDSR) and cfg2 lines on the end-of-round library
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 50 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu > $O/r4e_bench_cfg3_1gpu.json 2> $O/r4e_bench_cfg3_1gpu.err
tail -c 900 $O/r4e_bench_cfg3_1gpu.json | head -c 300; echo
timeout 70 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu > $O/r4e_bench_cfg2_1gpu.json 2> $O/r4e_bench_cfg2_1gpu.err
python - <<PY
import json
for w in ("cfg3", "cfg2"):
    try:
        j=json.loads(open("$O/r4e_bench_%s_1gpu.json" % w).read().strip().split("\n")[-1])
        print(w, round(j["value"],1), "ms", round(j["ms_per_step"],3), "e2e", j["e2e"] and round(j["e2e"]["value"],1), "frac", round(j["roofline"]["frac"],4))
    except Exception as e:
        print(w, "failed", e)
PY
echo done
