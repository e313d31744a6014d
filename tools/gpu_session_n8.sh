#!/bin/bash
# 8-GPU call: bench.py --gpus 8 as the driver launches it (512^3 strong-scaling line + BASELINE config 5, 2048^3)
tag=${1:-r02n8}
out=gpurun_out
mkdir -p $out
start=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 20 --warmup 3 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
echo "bench n8 rc=$? wall $(( $(date +%s) - start )) s"
grep -v "^$" $out/${tag}_bench_n8.err | tail -5
cut -c1-400 $out/${tag}_bench_n8.json
