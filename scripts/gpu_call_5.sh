#!/bin/bash
mkdir -p gpurun_out
AB_BUDGET_S=30 AB_ONLY="default,512 threads" python scripts/dev_sk_ab.py > gpurun_out/r02_sk_ab5_c2.txt 2>&1; echo "ab c2 rc=$?"
AB_BUDGET_S=45 AB_ONLY="default,512 threads" python scripts/dev_sk_ab.py 50000000 150 > gpurun_out/r02_sk_ab5_50m.txt 2>&1; echo "ab 50m rc=$?"
cut -c1-200 gpurun_out/r02_sk_ab5_c2.txt gpurun_out/r02_sk_ab5_50m.txt
