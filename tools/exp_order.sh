#!/bin/bash
# in-bin order x force-tile pitch experiment (tools/bench_pm_kernels.py, "local sort" lines)
out=gpurun_out
for v in "A=0" "PSC_SORT_KEY=cell"; do   # (the 16/160 force-tile pitch of the recorded experiment was removed)
  echo "== $v" >> $out/$1_exp_order.log
  env $v timeout 300 python tools/bench_pm_kernels.py 9 2>&1 | grep "local sort" >> $out/$1_exp_order.log
done
cat $out/$1_exp_order.log
