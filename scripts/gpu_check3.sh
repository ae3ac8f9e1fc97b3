#!/bin/bash
# 1-GPU pass: tests, localization of the count mismatch, warm-level sweep, scatter ncu.
set +e
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seed 1235 > gpurun_out/bench_n1_seed1235.json 2> gpurun_out/bench_n1_seed1235.err
echo "rc=$?" >> gpurun_out/bench_n1_seed1235.err
for W in 0 1048576 2097152 4194304 8388608; do
  BIGSLICE_WARM_SLOTS=$W timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n1_warm$W.json 2> gpurun_out/bench_n1_warm$W.err
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xchg_scatter -s 20 -c 1 -f -o gpurun_out/prof_scatter_v0b \
   python benchmarks/reshard_bw.py --max-mb 1024 --iters 1 --out gpurun_out/reshard_1gpu.jsonl > gpurun_out/ncu_scatter.log 2>&1
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
echo seed1235; tail -2 gpurun_out/bench_n1_seed1235.err; head -c 300 gpurun_out/bench_n1_seed1235.json; echo
for W in 0 1048576 2097152 4194304 8388608; do echo warm=$W; grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_n1_warm$W.json | head -2; done
cat gpurun_out/bench_n1.json; tail -8 gpurun_out/ncu_scatter.log
