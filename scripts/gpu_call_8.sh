#!/bin/bash
# round 2, 8-GPU call: bench with parity (weak C2, strong C3, C4 at full size, scaled C5), a bounded set of the sharded parity tests, a trace
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 \
    --c4 --c5 --c5-genome 100000000 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_n${N}.err
CDBG_SHARD_TEST_SUBSET=1 timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -rxXs -k "golden or c2_like or multi_word or abundance or even_k" > gpurun_out/r02_tests_sharded_${N}gpu.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_tests_sharded_${N}gpu.txt
CDBG_TRACE=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 scripts/dev_shard_trace.py \
    > gpurun_out/r02_trace_n${N}.txt 2>&1; echo "trace rc=$?"
ls -la gpurun_out
