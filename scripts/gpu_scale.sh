#!/bin/bash
# Scaling pass on N GPUs of one box: headline bench, reshard sweep, cogroup join.
# usage: bash scripts/gpu_scale.sh N
set +e
N=${1:-8}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" -eq 2 ]; then
  timeout 600 python -m pytest tests/test_spmd.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_multigpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_multigpu.log; tail -3 gpurun_out/pytest_multigpu.log
fi
timeout 120 python benchmarks/h2d_bw.py > gpurun_out/h2d_bw.json 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 900 bash -c "$(declare -f run); N=$N run 29571 bench.py --gpus $N --steps 10 --warmup 3" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?" >> gpurun_out/bench_n$N.err
timeout 900 bash -c "$(declare -f run); N=$N run 29572 benchmarks/reshard_bw.py --max-mb 1024 --out gpurun_out/reshard_bw_n$N.jsonl" > gpurun_out/reshard_bw_n$N.log 2>&1
timeout 900 bash -c "$(declare -f run); N=$N run 29573 benchmarks/cogroup_join.py --large 67108864 --small 4194304 --out gpurun_out/cogroup_n$N.json" > gpurun_out/cogroup_n$N.log 2>&1
timeout 900 bash -c "$(declare -f run); N=$N run 29574 bench.py --gpus $N --steps 5 --warmup 3 --shuffle nccl --no-e2e" > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
timeout 900 bash -c "$(declare -f run); N=$N run 29575 benchmarks/urls_topn.py --rows 134217728 --out gpurun_out/urls_topn_n$N.json" > gpurun_out/urls_topn_n$N.log 2>&1
tail -1 gpurun_out/urls_topn_n$N.log
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json; cat gpurun_out/bench_n${N}_nccl.json
tail -8 gpurun_out/reshard_bw_n$N.log; tail -2 gpurun_out/cogroup_n$N.log
