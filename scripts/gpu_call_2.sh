#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): sharded parity tests kept as a log, bench at N with parity + C3, trace of one sharded step
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -rxXs > gpurun_out/r02_tests_sharded_${N}gpu.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_tests_sharded_${N}gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02_bench_n${N}.err
CDBG_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 scripts/dev_shard_trace.py \
    > gpurun_out/r02_trace_n${N}.txt 2>&1; echo "trace rc=$?"
ls -la gpurun_out
