# usage: r02_scale_als.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 --workload als100m > gpurun_out/r02_als100m_n$N.json 2> gpurun_out/r02_als100m_n$N.log; echo als100m rc=$?
grep -E "\[scale\]|Error|Traceback" gpurun_out/r02_als100m_n$N.log | tail -20
cat gpurun_out/r02_als100m_n$N.json | head -c 600
