set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/ncu_r2
SMALL="--systems-total 16 --step-systems 16 --max-iter 100 --no-extras"
# 1. launch list of a (shortened) bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/ncu_r2/launches_r2.csv python bench.py --steps 2 --warmup 1 $SMALL > gpurun_out/ncu_r2/bench_under_ncu.log 2>&1
# 2. full capture of the fused kernel (16 systems, 40 bodies)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pcg_fused -c 1 -f -o gpurun_out/ncu_r2/fused_r2 python bench.py --steps 1 --warmup 1 --systems-total 16 --step-systems 16 --max-iter 40 --no-extras > gpurun_out/ncu_r2/fused.log 2>&1
# 3. tile-stream solve, 8 x 256^3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sptrsv_ts -s 2 -c 1 -f -o gpurun_out/ncu_r2/ts_r2 python tools/gpu_trsv_ts.py --sides 256 --batches 8 --position-space 1 --syncfree 0 > gpurun_out/ncu_r2/ts.log 2>&1
# 4. level-stream solve, one 316^2 factor
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sptrsv_ls -s 1 -c 1 -f -o gpurun_out/ncu_r2/ls_r2 python tools/profile_ls.py > gpurun_out/ncu_r2/ls.log 2>&1
ls -la gpurun_out/ncu_r2
tail -3 gpurun_out/ncu_r2/*.log
