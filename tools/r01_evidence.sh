set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r01k_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01k_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py > gpurun_out/r01k_bench_n1.json 2> gpurun_out/r01k_bench_n1.log; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01k_bench_ref.json 2> gpurun_out/r01k_bench_ref.log; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --profile --variants bf16 > gpurun_out/r01k_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel -c 2 -f -o gpurun_out/r01k_als_tc python bench.py --steps 1 --warmup 0 --no-cpu --no-knn --profile --variants bf16 > gpurun_out/r01k_ncu_als.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_build_kernel -c 1 -f -o gpurun_out/r01k_knn python bench.py --steps 1 --warmup 0 --no-cpu --profile --variants bf16 > gpurun_out/r01k_ncu_knn.log 2>&1
tail -3 gpurun_out/r01k_pytest.log; cat gpurun_out/r01k_bench_n1.json; cat gpurun_out/r01k_bench_ref.json
