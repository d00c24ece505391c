#!/bin/bash
mkdir -p gpurun_out
AB_BUDGET_S=40 AB_ONLY="k-mer instances,default,two-phase,K6 full" python scripts/dev_sk_ab.py > gpurun_out/r02_sk_ab4_c2.txt 2>&1; echo "ab c2 rc=$?"
AB_BUDGET_S=60 AB_ONLY="default,two-phase insert,K6 full" python scripts/dev_sk_ab.py 50000000 150 > gpurun_out/r02_sk_ab4_50m.txt 2>&1; echo "ab 50m rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_quick2.json 2>&1; echo "bench rc=$?"
