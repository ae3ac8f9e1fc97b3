#!/bin/bash
for R in init clear init clear; do
for S in 1 2; do
  BIGSLICE_TABLE_RECYCLE=$R timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --shards-per-gpu $S > gpurun_out/b_$R$S.json 2> gpurun_out/b_$R$S.err
  echo "recycle=$R shards=$S: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/b_$R$S.json | head -1)"; tail -1 gpurun_out/b_$R$S.err | cut -c1-200
done; done
for R in init clear; do
BIGSLICE_TABLE_RECYCLE=$R timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --shards-per-gpu 1 > gpurun_out/ncu_launches_$R.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_$R.csv | grep -v "at::\|at_cuda"
done
