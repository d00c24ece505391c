#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02c_bench_n2.err
timeout 300 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_cli.py -m gpu -x -q -rxXs -k "failure_does_not_hang or multi_gpu or c2_like" > gpurun_out/r02c_tests_2gpu.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02c_tests_2gpu.txt
