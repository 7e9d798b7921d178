set -x
mkdir -p gpurun_out
timeout 300 python tools/score_prof.py 2048 > gpurun_out/r02f_score_plain.log 2>&1; tail -3 gpurun_out/r02f_score_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_score_dense -c 1 -f -o gpurun_out/r02f_score_dense python tools/score_prof.py 2048 > gpurun_out/r02f_ncu_score.log 2>&1; tail -2 gpurun_out/r02f_ncu_score.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tcx_kernel -c 2 -f -o gpurun_out/r02f_als_tcx python bench.py --steps 1 --warmup 0 --no-cpu --no-knn --profile --variants fp32 > gpurun_out/r02f_ncu_tcx.log 2>&1; tail -2 gpurun_out/r02f_ncu_tcx.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel -c 2 -f -o gpurun_out/r02f_als_tc python bench.py --steps 1 --warmup 0 --no-cpu --no-knn --profile --variants bf16 > gpurun_out/r02f_ncu_tc.log 2>&1; tail -2 gpurun_out/r02f_ncu_tc.log
ls -la gpurun_out/*.ncu-rep
