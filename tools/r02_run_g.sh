mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_als_tc128_gpu.py tests/test_knn_gpu.py tests/test_user_knn_gpu.py tests/test_als_tc_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/r02g_pytest.log; tail -12 gpurun_out/r02g_pytest.log
timeout 300 python tools/score_prof.py 4096 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 2 --workload als100m --scale 0.05 > gpurun_out/r02g_als100m_n1_s05.json 2> gpurun_out/r02g_als100m_n1_s05.log; echo als100m rc=$?; grep -E "\[scale\]|Error" gpurun_out/r02g_als100m_n1_s05.log | tail
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.log; echo bench rc=$?; grep "\[bench\]" gpurun_out/r02g_bench.log | tail -30
