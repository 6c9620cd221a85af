# Round 2: is cfg1 bound by the device or by the launching thread now that programmatic dependent launch hides the set-up?
# (wowsr_get_timing phase 4 = host milliseconds spent enqueuing a forward's launches, bench: roofline.host_enqueue_ms_per_step)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python bench.py --workload cfg1 --steps 30 --warmup 5 --no-cpu > $O/r4c_bench_cfg1.json 2> $O/r4c_bench_cfg1.err
python - <<PY
import json
j=json.loads(open("$O/r4c_bench_cfg1.json").read().strip().split("\n")[-1])
r=j["roofline"]
print("cfg1 step", round(j["ms_per_step"],3), "conv", round(r["conv_ms_per_step"],3), "host enqueue", round(r["host_enqueue_ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3))
PY
timeout 100 python -m pytest tests/test_abi.py tests/test_gpu_rrdbnet.py -m gpu -x -q > $O/r4c_pytest.txt 2>&1; echo "exit $?" >> $O/r4c_pytest.txt
tail -3 $O/r4c_pytest.txt
echo done
