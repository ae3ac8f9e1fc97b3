#!/bin/bash
for cfg in "init 1" "reinit 1" "reinit 2" "init 1"; do set -- $cfg
  BIGSLICE_TABLE_RECYCLE=$1 BIGSLICE_POOL_TABLES=$2 BIGSLICE_SPANS=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e > gpurun_out/b_$1$2.json 2> gpurun_out/b_$1$2.err
  echo "recycle=$1 tables=$2: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/b_$1$2.json | head -1)"
done
grep -A14 "timeline" gpurun_out/b_reinit1.err
