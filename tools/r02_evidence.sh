# Round-2 evidence on one B200: full GPU test suite, smoke, default bench, ncu launch list, one `--set full`
# capture per hot kernel, reference arm.  Outputs under gpurun_out/r02z_*; summaries go to profiles/.
set -x
mkdir -p gpurun_out
export LK_BENCH_DATA_CACHE=/tmp/lk_ml25m_cache.npz   # the same synthetic matrix for every invocation below
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02z_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02z_smoke.log
timeout 900 python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.log; echo "bench rc=$?"
head -c 2500 gpurun_out/r02z_bench_n1.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --profile > gpurun_out/r02z_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel -c 2 -f -o gpurun_out/r02z_als_tc python bench.py --steps 1 --warmup 0 --no-cpu --no-knn --profile --variants bf16 > gpurun_out/r02z_ncu_als.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:knn_build_kernel -c 1 -f -o gpurun_out/r02z_knn python bench.py --steps 1 --warmup 0 --no-cpu --profile --variants bf16 > gpurun_out/r02z_ncu_knn.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:als_tcx_kernel -c 1 -f -o gpurun_out/r02z_als_tcx python bench.py --steps 1 --warmup 0 --no-cpu --no-knn --profile --variants fp32 > gpurun_out/r02z_ncu_tcx.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:als_tc128_kernel -c 1 -f -o gpurun_out/r02z_als_tc128 python bench.py --workload als100m --scale 0.05 --steps 1 --warmup 1 --variants bf16 --no-parity > gpurun_out/r02z_ncu_tc128.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:knn_score_dense -c 1 -f -o gpurun_out/r02z_score_dense python tools/score_prof.py 2048 > gpurun_out/r02z_ncu_score.log 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.log; echo "ref rc=$?"
head -c 1200 gpurun_out/r02z_bench_ref.json; echo
ls -la gpurun_out/r02z_*
