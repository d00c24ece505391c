#!/bin/bash
# 8-GPU bench only (the command the driver runs), with the stage trace afterwards
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 \
    > gpurun_out/r02d_bench_n8.json 2> gpurun_out/r02d_bench_n8.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02d_bench_n8.err
CDBG_TRACE=1 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 scripts/dev_shard_trace.py \
    > gpurun_out/r02d_trace_n8.txt 2>&1; echo "trace rc=$?"
