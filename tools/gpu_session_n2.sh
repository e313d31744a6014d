#!/bin/bash
# 2-GPU call: slab path over NCCL against the single-domain path (FFT, multigrid, QUMOND, f(R)), then bench.py --gpus 2
tag=${1:-r02n2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cfg in "7 fft newton" "6 multigrid newton" "6 fft_7pt mond" "6 multigrid fr"; do
  timeout 300 $TR --master-port 29517 tools/check_slab_multigpu.py $cfg > $out/${tag}_check_$(echo $cfg | tr ' ' '_').log 2>&1
  echo "check $cfg rc=$?"
  tail -4 $out/${tag}_check_$(echo $cfg | tr ' ' '_').log
done
timeout 900 $TR --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench512_n2.json 2> $out/${tag}_bench512_n2.err
echo "bench n2 rc=$?"
tail -3 $out/${tag}_bench512_n2.err
cut -c1-600 $out/${tag}_bench512_n2.json
