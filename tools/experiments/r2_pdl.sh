# Round 2: programmatic dependent launch of the rolling kernel (option pdl; build flag WOWSR_PDL_DEFAULT).
# A/B on one box through the SAME library (--opt pdl=0/1): cfg1 (354 dependent launches of 7-17 us), cfg2, cfg5 scene;
# then the whole GPU suite on the library built with the default ON (build/libwowsr_pdl1.so).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
T0=$SECONDS
for p in 0 1; do
  timeout 200 python bench.py --workload cfg1 --steps 30 --warmup 5 --no-cpu --opt pdl=$p > $O/r4a_bench_cfg1_pdl$p.json 2> $O/r4a_bench_cfg1_pdl$p.err
done
for p in 0 1; do
  timeout 200 python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu --no-e2e --opt pdl=$p > $O/r4a_bench_cfg2_pdl$p.json 2> $O/r4a_bench_cfg2_pdl$p.err
done
python - <<PY
import json
for w in ("cfg1", "cfg2"):
    for p in (0, 1):
        try:
            j = json.loads(open("$O/r4a_bench_%s_pdl%d.json" % (w, p)).read().strip().split("\n")[-1])
            print(w, "pdl", p, "ms", round(j["ms_per_step"], 3), "e2e", j.get("e2e") and round(j["e2e"]["ms_per_step"], 3), "launches", j["gpu_launches"], j["clocks"])
        except Exception as e:
            print(w, p, "failed", e)
PY
WOWSR_LIB=$GRAFT_REPO_ROOT/build/libwowsr_pdl1.so timeout 900 python -m pytest tests -m gpu -x -q > $O/r4a_pytest_gpu_pdl1.txt 2>&1; echo "exit $?" >> $O/r4a_pytest_gpu_pdl1.txt
tail -4 $O/r4a_pytest_gpu_pdl1.txt
[ $((SECONDS - T0)) -gt 400 ] && { echo "skipping the scene A/B"; echo done; exit 0; }
for p in 0 1; do
  timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --opt pdl=$p > $O/r4a_bench_scene_pdl$p.json 2> $O/r4a_bench_scene_pdl$p.err
  tail -c 1200 $O/r4a_bench_scene_pdl$p.json | head -c 400; echo
done
echo done
