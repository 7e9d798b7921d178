mkdir -p gpurun_out
for f in 0 1 2 3; do echo "== LK_ALS_FLAGS=$f"; LK_ALS_FLAGS=$f python tools/als_phase_prof.py 2>&1 | tail -22; done > gpurun_out/r02_e1_phase.log 2>&1
cat gpurun_out/r02_e1_phase.log | grep -E "==|half|solve4|gather|preload|write"
python -m pytest tests/test_als_gpu.py tests/test_scale_parity.py -x -q 2>&1 | tail -3
