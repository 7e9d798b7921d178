# usage: r02_bench_n.sh N — the default bench line on N GPUs (what the driver runs for SCALE)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.log; echo rc=$?
grep -E "\[bench\]|Error|Traceback" gpurun_out/r02_bench_n$N.log | tail -12
head -c 600 gpurun_out/r02_bench_n$N.json
