# usage: r02_scale_n8.sh N — the two scale configurations and one default bench line on N GPUs
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29532 bench.py --gpus $N --steps 3 --warmup 3 --workload knn1b > gpurun_out/r02_knn1b_n$N.json 2> gpurun_out/r02_knn1b_n$N.log; echo knn1b rc=$?
grep -E "\[scale\]|Error|Traceback" gpurun_out/r02_knn1b_n$N.log | tail -8
$TR --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 --workload als100m > gpurun_out/r02_als100m_n$N.json 2> gpurun_out/r02_als100m_n$N.log; echo als100m rc=$?
grep -E "\[scale\]|Error|Traceback" gpurun_out/r02_als100m_n$N.log | tail -12
$TR --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.log; echo bench rc=$?
tail -5 gpurun_out/r02_bench_n$N.log
head -c 1500 gpurun_out/r02_bench_n$N.json
