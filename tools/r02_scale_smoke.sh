mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --workload als100m --scale 0.02 > gpurun_out/r02d_als100m_smoke.json 2> gpurun_out/r02d_als100m_smoke.log; echo als rc=$?
grep -E "\[scale\]|Error|error|Traceback" gpurun_out/r02d_als100m_smoke.log | tail -20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --workload knn1b --scale 0.01 > gpurun_out/r02d_knn1b_smoke.json 2> gpurun_out/r02d_knn1b_smoke.log; echo knn rc=$?
grep -E "\[scale\]|Error|error|Traceback" gpurun_out/r02d_knn1b_smoke.log | tail -20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_check.py > gpurun_out/r02d_mgpu_check.log 2>&1; echo mgpu rc=$?; tail -5 gpurun_out/r02d_mgpu_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.log; echo bench2 rc=$?
grep -E "\[bench\]|Error|Traceback" gpurun_out/r02d_bench_n2.log | tail -20
