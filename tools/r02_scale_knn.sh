# usage: r02_scale_knn.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 2 --warmup 1 --workload knn1b > gpurun_out/r02_knn1b_n$N.json 2> gpurun_out/r02_knn1b_n$N.log; echo knn1b rc=$?
grep -E "\[scale\]|Error|Traceback" gpurun_out/r02_knn1b_n$N.log | tail -20
cat gpurun_out/r02_knn1b_n$N.json | head -c 600
