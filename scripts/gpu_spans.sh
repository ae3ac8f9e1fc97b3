#!/bin/bash
set +e
mkdir -p gpurun_out
BIGSLICE_SPANS=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e > gpurun_out/spans_n1.json 2> gpurun_out/spans_n1.err
BIGSLICE_SPANS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/spans_n2.json 2> gpurun_out/spans_n2.err
grep -A45 "host spans" gpurun_out/spans_n1.err | grep -v "^\[\|Warning"; head -c 300 gpurun_out/spans_n1.json; echo
grep -A60 "host spans" gpurun_out/spans_n2.err | grep -v "^\[\|Warning"; head -c 300 gpurun_out/spans_n2.json; echo
