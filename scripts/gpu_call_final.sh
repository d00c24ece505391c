#!/bin/bash
# round 2, final single-GPU artefacts: full -m gpu suite, bench (both arms), launch list, full ncu captures of the kernels that changed
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rxXs > gpurun_out/r02e_tests_1gpu.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02e_tests_1gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02e_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-to-file --no-c3 --no-parity > gpurun_out/r02e_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sk1_extract|sk2p_hist|sk2p_scatter|sk2c_count|k3_ends|k5p_scatter|k5c_join|k6_local|k6_resolve" -c 10 -f -o gpurun_out/r02e_full_a \
    python scripts/dev_c2.py 1.0 1 > gpurun_out/r02e_full_a.log 2>&1; echo "ncu full a rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k6_jump_list|k8_spell" -c 5 -f -o gpurun_out/r02e_full_b \
    python scripts/dev_c2.py 1.0 1 > gpurun_out/r02e_full_b.log 2>&1; echo "ncu full b rc=$?"
ls -la gpurun_out | tail -12
