#!/bin/bash
# round 2, first single-GPU call: tests, bench (both arms), launch list, full ncu captures of the top kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rxXs > gpurun_out/r02_tests_1gpu.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_tests_1gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-to-file --no-c3 --no-parity > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sk1_extract|sk2p_hist|sk2p_scatter|sk2c_count|k3_ends|k5p_hist|k5p_scatter|k5c_join" -c 9 -f -o gpurun_out/r02_full_a \
    python scripts/dev_c2.py 1.0 1 > gpurun_out/r02_full_a.log 2>&1; echo "ncu full a rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k6_jump" --launch-skip 1 -c 2 -f -o gpurun_out/r02_full_b \
    python scripts/dev_c2.py 1.0 1 > gpurun_out/r02_full_b.log 2>&1; echo "ncu full b rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k8_spell|k6_finalize|k9_" -c 4 -f -o gpurun_out/r02_full_c \
    python scripts/dev_c2.py 1.0 1 > gpurun_out/r02_full_c.log 2>&1; echo "ncu full c rc=$?"
ls -la gpurun_out
