mkdir -p gpurun_out
for W in 16 8; do
  LK_KNN_WARPS=$W timeout 400 python bench.py --steps 2 --warmup 1 --workload knn1b --scale 0.1 --no-parity > gpurun_out/r02_knn_sweep_w$W.json 2> gpurun_out/r02_knn_sweep_w$W.log; echo "W=$W rc=$?"
  grep -E "\[scale\] knn1b x" gpurun_out/r02_knn_sweep_w$W.log
done
LK_KNN_WARPS=8 LK_KNN_CTAS=3 timeout 400 python bench.py --steps 2 --warmup 1 --workload knn1b --scale 0.1 --no-parity 2>&1 | grep -E "\[scale\] knn1b x"
