#!/bin/bash
mkdir -p gpurun_out
AB_BUDGET_S=70 python scripts/dev_sk_ab.py > gpurun_out/r02_sk_ab3_c2.txt 2>&1; echo "ab c2 rc=$?"
AB_BUDGET_S=100 AB_ONLY="k-mer instances,default,dedup,two-level" python scripts/dev_sk_ab.py 50000000 150 > gpurun_out/r02_sk_ab3_50m.txt 2>&1; echo "ab 50m rc=$?"
python -m pytest tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/r02_tests_cli.txt 2>&1; echo "cli tests rc=$?"; tail -3 gpurun_out/r02_tests_cli.txt
python bench.py --steps 5 --warmup 3 --no-c3 --no-cpu-baseline > gpurun_out/r02_bench_quick.json 2>&1; echo "bench rc=$?"
