#!/bin/bash
# One gpurun call: GPU tests, kernel timings, the bench line, ncu (launch list + full captures of the particle kernels
# + every grid kernel).   usage: gpurun --timeout 3000 -- 'bash tools/gpu_session.sh r02b'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 600 python tools/bench_pm_kernels.py 9 > $out/${tag}_pm_kernels.log 2>&1
if [ -z "$SKIP_MG" ]; then
timeout 600 python tools/bench_multigrid.py 9 > $out/${tag}_multigrid512.log 2>&1
timeout 600 python tools/bench_multigrid.py 8 > $out/${tag}_multigrid256.log 2>&1
fi
timeout 1200 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench512.json 2> $out/${tag}_bench512.err
echo "bench rc=$?" >> $out/${tag}_bench512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > $out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'deposit_binned_kernel|interp_kick_phi_binned_kernel|step_sort' \
  --launch-skip 8 --launch-count 8 -o $out/${tag}_particle_kernels -f python tools/prof_step.py 9 step > $out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py $out/${tag}_particle_kernels.ncu-rep > $out/${tag}_particle_kernels_ncu.txt 2>&1
if [ -n "$REFARM" ]; then
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_reference_arm.json 2> $out/${tag}_reference_arm.err
echo "reference arm rc=$?" >> $out/${tag}_reference_arm.err
fi
if [ -n "$DIAG3" ]; then
for v in "" "PSC_NO_FUSED_GS=1" "PSC_ORDER=reference"; do env $v timeout 300 python tools/diag_config3.py 8 14 >> $out/${tag}_diag3.log 2>&1; done
fi
if [ -z "$SKIP_GRID_NCU" ]; then bash tools/gpu_session2.sh $tag; fi
ls -la $out | tail -30
tail -5 $out/${tag}_pytest.log
cat $out/${tag}_pm_kernels.log
tail -12 $out/${tag}_multigrid512.log 2>/dev/null
