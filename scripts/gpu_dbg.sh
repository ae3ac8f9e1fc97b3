#!/bin/bash
timeout 900 python -m pytest tests/test_spmd.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -3
DBG=base timeout 300 python scripts/debug_table2.py 2>&1 | tail -2
for S in 1 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --shards-per-gpu $S > gpurun_out/b_$S.json 2> gpurun_out/b_$S.err
  echo "shards=$S: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/b_$S.json | head -1)"; tail -1 gpurun_out/b_$S.err | cut -c1-200
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_2.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --shards-per-gpu 2 > gpurun_out/ncu_launches_2.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_2.csv | grep -v "at::\|at_cuda"
