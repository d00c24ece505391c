#!/bin/bash
# 4 GPUs: the driver's bench command (weak C2 + strong C3, parity) -- the one world size the builder had not run
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 20 --warmup 5 \
    > gpurun_out/r02g_bench_n4.json 2> gpurun_out/r02g_bench_n4.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02g_bench_n4.err
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide_9 or circular or pow2 or reads_3" > gpurun_out/r02g_tests_rank.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02g_tests_rank.txt
