#!/bin/bash
# First GPU pass: numerics tests, smoke, short bench, launch list + one full ncu capture.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --rows-per-gpu $((1<<25)) > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hash_combine_tma -s 3 -c 1 -f -o gpurun_out/prof_combine \
   python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
