#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
BIGSLICE_SPANS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err
grep -A40 "host spans" gpurun_out/bench_n${N}_final.err | grep -v "^\[\|Warning"; cat gpurun_out/bench_n${N}_final.json
