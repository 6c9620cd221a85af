# Round 2, closing call: whole GPU suite + smoke + bench lines (cfg1, default scene) on the committed state (programmatic
# dependent launch on by default)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > $O/r4b_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r4b_pytest_gpu.txt
tail -4 $O/r4b_pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke > $O/r4b_smoke.txt 2>&1; tail -1 $O/r4b_smoke.txt
timeout 200 python bench.py --workload cfg1 --steps 30 --warmup 5 > $O/r4b_bench_cfg1_1gpu.json 2> $O/r4b_bench_cfg1_1gpu.err
tail -c 2500 $O/r4b_bench_cfg1_1gpu.json | head -c 600; echo
[ $((SECONDS - T0)) -gt 330 ] && { echo "skipping the scene line"; echo done; exit 0; }
timeout 400 python bench.py --steps 2 --warmup 3 > $O/r4b_bench_scene_1gpu.json 2> $O/r4b_bench_scene_1gpu.err
python - <<PY
import json
j=json.loads(open("$O/r4b_bench_scene_1gpu.json").read().strip().split("\n")[-1])
print("scene", round(j["value"],1), "ms", round(j["ms_per_step"],1), "e2e", round(j["e2e"]["value"],1), "roofline", round(j["roofline"]["frac"],4), j["clocks"])
PY
echo done
