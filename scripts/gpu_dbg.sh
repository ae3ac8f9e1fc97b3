#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
for S in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --shards-per-gpu $S > gpurun_out/b_$S.json 2> gpurun_out/b_$S.err
  echo "shards=$S: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/b_$S.json | head -1)"; tail -1 gpurun_out/b_$S.err | cut -c1-200
done
BIGSLICE_SPANS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/spans_n2.json 2> gpurun_out/spans_n2.err
grep -A40 "host spans" gpurun_out/spans_n2.err | grep -v "^\[\|Warning"; cat gpurun_out/spans_n2.json
