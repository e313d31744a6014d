#!/bin/bash
# 4-GPU call: bench.py --gpus 4 as the driver launches it
tag=${1:-r02n4}
out=gpurun_out
mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 4 --steps 20 --warmup 3 > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.err
echo "bench n4 rc=$?"
grep -v "^$" $out/${tag}_bench_n4.err | tail -3
cut -c1-300 $out/${tag}_bench_n4.json
