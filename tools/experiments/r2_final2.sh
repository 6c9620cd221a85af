# Round 2, last call: whole GPU suite + smoke on the committed state
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r3y_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r3y_pytest_gpu.txt
tail -4 $O/r3y_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/r3y_smoke.txt 2>&1; tail -1 $O/r3y_smoke.txt
timeout 600 python bench.py --steps 2 --warmup 3 > $O/r3y_bench_scene_1gpu.json 2> $O/r3y_bench_scene_1gpu.err
python - <<PY
import json
j=json.loads(open("$O/r3y_bench_scene_1gpu.json").read().strip().split("\n")[-1])
print("scene", round(j["value"],1), "ms", round(j["ms_per_step"],1), "e2e", round(j["e2e"]["value"],1), "roofline", round(j["roofline"]["frac"],4), j["clocks"])
PY
echo done
