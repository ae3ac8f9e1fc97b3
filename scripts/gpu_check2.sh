#!/bin/bash
# Second GPU pass (2 GPUs): full GPU test suite incl. multi-GPU SPMD, bench at N=1,2,
# ncu of the reworked combine kernel, reshard bandwidth sweep, cogroup join.
set +e
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG" > gpurun_out/gpu2.txt
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hash_combine_tma -s 3 -c 1 -f -o gpurun_out/prof_combine_v2b \
   python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xchg_scatter -s 8 -c 1 -f -o gpurun_out/prof_scatter_v0 \
   python benchmarks/reshard_bw.py --max-mb 256 --iters 2 --out gpurun_out/reshard_1gpu.jsonl > gpurun_out/ncu_scatter.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seed 1235 > gpurun_out/bench_n1_seed1235.json 2> gpurun_out/bench_n1_seed1235.err
echo "rc=$?" >> gpurun_out/bench_n1_seed1235.err
for W in 0 1048576 2097152 8388608; do
  BIGSLICE_WARM_SLOTS=$W timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n1_warm$W.json 2> gpurun_out/bench_n1_warm$W.err
done
if [ "$NG" -ge 2 ]; then
  BIGSLICE_WARM_SLOTS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 \
     bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n2_warm0.json 2> gpurun_out/bench_n2_warm0.err
  echo "rc=$?" >> gpurun_out/bench_n2_warm0.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
     bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
  echo "bench2 rc=$?" >> gpurun_out/bench_n2.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
     benchmarks/reshard_bw.py --max-mb 1024 > gpurun_out/reshard_bw.log 2>&1
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 \
     benchmarks/cogroup_join.py --large 33554432 --small 2097152 > gpurun_out/cogroup.log 2>&1
fi
cat gpurun_out/debug_local.log
tail -4 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_n1.err gpurun_out/bench_n2.err
cat gpurun_out/bench_n1.json gpurun_out/bench_n2.json
for f in gpurun_out/bench_n1_seed1235 gpurun_out/bench_n2_warm0; do echo $f; tail -2 $f.err; cat $f.json | head -c 300; echo; done
for W in 0 1048576 2097152 8388608; do echo warm=$W; head -c 250 gpurun_out/bench_n1_warm$W.json; echo; done
tail -12 gpurun_out/reshard_bw.log
tail -3 gpurun_out/cogroup.log
