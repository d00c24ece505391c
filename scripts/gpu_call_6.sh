#!/bin/bash
mkdir -p gpurun_out
AB_BUDGET_S=20 AB_ONLY="default,small staging" python scripts/dev_sk_ab.py > gpurun_out/r02_sk_ab6_c2.txt 2>&1
AB_BUDGET_S=30 AB_ONLY="default,small staging" python scripts/dev_sk_ab.py 50000000 150 > gpurun_out/r02_sk_ab6_50m.txt 2>&1
cut -c1-175 gpurun_out/r02_sk_ab6_c2.txt gpurun_out/r02_sk_ab6_50m.txt
