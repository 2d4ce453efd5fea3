#!/bin/bash
# usage: r02_multi.sh <N> <tag> <bench args...>   -- one bench line on N GPUs of this box (torchrun for N > 1)
N=$1; TAG=$2; shift 2
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 "$@" > gpurun_out/r02_bench_${TAG}_${N}gpu.json 2> gpurun_out/r02_bench_${TAG}_${N}gpu.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N "$@" > gpurun_out/r02_bench_${TAG}_${N}gpu.json 2> gpurun_out/r02_bench_${TAG}_${N}gpu.err
fi
echo "rc=$?" >> gpurun_out/r02_bench_${TAG}_${N}gpu.err
tail -4 gpurun_out/r02_bench_${TAG}_${N}gpu.err | cut -c1-300
head -c 1500 gpurun_out/r02_bench_${TAG}_${N}gpu.json
