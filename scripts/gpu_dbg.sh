#!/bin/bash
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_default.json)"; head -c 250 gpurun_out/bench_default.json; echo
