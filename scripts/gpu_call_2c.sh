#!/bin/bash
mkdir -p gpurun_out
CDBG_TRACE=1 DEV_G=250000000 DEV_L=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 scripts/dev_shard_trace.py \
    > gpurun_out/r02c_trace_c3_n2.txt 2>&1; echo "trace rc=$?"
timeout 120 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -rxXs -k "failure_does_not_hang" > gpurun_out/r02c_tests_fault.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02c_tests_fault.txt
timeout 200 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "multi_gpu" > gpurun_out/r02c_tests_cli_nbgpus.txt 2>&1; echo "cli rc=$?"; tail -3 gpurun_out/r02c_tests_cli_nbgpus.txt
