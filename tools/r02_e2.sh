mkdir -p gpurun_out
tools/_bin/chol_tc_bench > gpurun_out/r02_chol_tc_bench.log 2>&1
grep -E "occ|phases|per step|warp 3" gpurun_out/r02_chol_tc_bench.log | cut -c1-330
for v in "LK_ALS_GJ=0 LK_ALS_FLAGS=1" "LK_ALS_GJ=0 LK_ALS_FLAGS=0" "LK_ALS_GJ=1 LK_ALS_FLAGS=1"; do
  echo "== $v"
  env $v python bench.py --steps 30 --warmup 5 --variants bf16 --no-knn --no-cpu 2>gpurun_out/e2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d.get('parity',{}).get('als_bf16',{}).get('ok'))"
done
