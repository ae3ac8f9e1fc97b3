#!/bin/bash
# 1-GPU profiling pass: bench (1 and 2 shards per GPU), launch lists under ncu, full
# ncu capture of the combine and the exchange-scatter kernels.
set +e
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --shards-per-gpu 2 > gpurun_out/bench_n1_s2.json 2> gpurun_out/bench_n1_s2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s2.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --shards-per-gpu 2 > gpurun_out/ncu_launches_s2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s1.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_launches_s1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hash_combine_tma -s 3 -c 1 -f -o gpurun_out/prof_combine_v3 \
   python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xchg_scatter -s 3 -c 1 -f -o gpurun_out/prof_scatter_table \
   python bench.py --steps 1 --warmup 3 --no-e2e --shards-per-gpu 2 > gpurun_out/ncu_full2.log 2>&1
head -c 400 gpurun_out/bench_n1.json; echo; head -c 400 gpurun_out/bench_n1_s2.json; echo
python scripts/summarize_launches.py gpurun_out/launches_s2.csv 
python scripts/summarize_launches.py gpurun_out/launches_s1.csv
ls -la gpurun_out/*.ncu-rep
