# Round 2: measured-speed balancing, headline workload (cfg5 scene) A/B on the same box
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
for a in 0 1 0 1; do
  timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --opt roll_adapt=$a > $O/r3n_bench_scene_adapt$a.json 2> $O/r3n_bench_scene_adapt$a.err
  echo adapt=$a $(grep -o '"ms_per_step": [0-9.]*' $O/r3n_bench_scene_adapt$a.json) $(grep -o '"sm_mhz": [0-9.]*' $O/r3n_bench_scene_adapt$a.json) $(grep -o '"power_w": [0-9.]*' $O/r3n_bench_scene_adapt$a.json)
done
echo done
