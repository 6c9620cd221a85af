# Round 2, post-process call: full ncu capture of the strip-march kernel (stall reasons, bank conflicts, pipes).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"post_march|clahe_hist" -c 2 -o $O/r2u_prof_post \
    python bench.py --workload post4096 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2u_ncu_run.log 2>&1
echo done
