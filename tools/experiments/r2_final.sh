# Round 2, final call: whole GPU suite + smoke on the committed state, bench lines of every workload, ncu --set full of the post-process kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r3z_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r3z_pytest_gpu.txt
tail -4 $O/r3z_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/r3z_smoke.txt 2>&1; tail -1 $O/r3z_smoke.txt
timeout 900 python bench.py > $O/r3z_bench_scene_1gpu.json 2> $O/r3z_bench_scene_1gpu.err
timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 > $O/r3z_bench_cfg2_1gpu.json 2> $O/r3z_bench_cfg2_1gpu.err
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 > $O/r3z_bench_cfg1_1gpu.json 2> $O/r3z_bench_cfg1_1gpu.err
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 5 > $O/r3z_bench_cfg3_1gpu.json 2> $O/r3z_bench_cfg3_1gpu.err
timeout 300 python bench.py --workload post4096 --steps 30 --warmup 5 > $O/r3z_bench_post4096_1gpu.json 2> $O/r3z_bench_post4096_1gpu.err
for f in scene cfg2 cfg1 cfg3 post4096; do python - <<PY
import json
j=json.loads(open("$O/r3z_bench_${f}_1gpu.json").read().strip().split("\n")[-1])
print("$f", round(j["value"],1), j["unit"], "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],1), "roofline", round(j["roofline"]["frac"],4), j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"post_march|clahe_hist" -c 2 -o $O/r3z_prof_post \
    python bench.py --workload post4096 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r3z_ncu_run.log 2>&1
echo done
