mkdir -p gpurun_out
python -m pytest tests/test_als_tcx_gpu.py tests/test_als_tc128_gpu.py tests/test_als_tc_gpu.py -x -q 2>&1 | tail -8
python bench.py --steps 30 --warmup 5 --no-knn --no-cpu 2>gpurun_out/e3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16', d['value'], d['roofline']['frac'], 'fp32', d['als_fp32']['ms_per_epoch'], d['parity']['als_bf16']['ok'], d['parity']['als_fp32']['ok'], d['parity']['als_fp32']['item']['rel_fro_vs_f64_oracle'])"
python bench.py --workload als100m --scale 0.05 --steps 5 --warmup 3 2>&1 | grep -E "\[scale\]|Error|error" | tail
