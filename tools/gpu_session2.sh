#!/bin/bash
# ncu --set full of every non-particle kernel (second launch of each), summarised into gpurun_out/<tag>_grid_kernels_ncu.txt
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'psc::' \
  -o $out/${tag}_grid_kernels -f python tools/prof_all_kernels.py 9 > $out/${tag}_ncu_grid.log 2>&1
python tools/ncu_summary.py $out/${tag}_grid_kernels.ncu-rep > $out/${tag}_grid_kernels_ncu_all.txt 2>&1
tail -3 $out/${tag}_ncu_grid.log
ls -la $out/${tag}_grid_kernels*
# the report itself is too large to travel (64 MiB limit of gpurun_out): the summary is what is kept
if [ $(stat -c %s $out/${tag}_grid_kernels.ncu-rep) -gt 30000000 ]; then rm $out/${tag}_grid_kernels.ncu-rep; fi
