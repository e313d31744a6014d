#!/bin/bash
# copy the evidence of a tools/gpu_session.sh run (gpurun_out/<tag>_*) into profiles/ under the round's names
# usage: tools/collect_profiles.sh r02m
tag=$1
g=gpurun_out
p=profiles
tail -1 $g/${tag}_bench512.json > $p/r02_bench512_n1.json
cp $g/${tag}_launches.csv $p/r02_launches_bench512.csv
cp $g/${tag}_particle_kernels_ncu.txt $p/r02_particle_kernels_ncu.txt
cp $g/${tag}_pm_kernels.log $p/r02_pm_kernels_512.txt
[ -f $g/${tag}_multigrid512.log ] && cp $g/${tag}_multigrid512.log $p/r02_multigrid_512.txt
[ -f $g/${tag}_multigrid256.log ] && cp $g/${tag}_multigrid256.log $p/r02_multigrid_256.txt
[ -f $g/${tag}_reference_arm.json ] && tail -1 $g/${tag}_reference_arm.json > $p/r02_bench_reference_arm.json
[ -f $g/${tag}_grid_kernels_ncu_all.txt ] && cp $g/${tag}_grid_kernels_ncu_all.txt $p/r02_grid_kernels_ncu.txt
[ -f $g/${tag}_pytest.log ] && tail -3 $g/${tag}_pytest.log > $p/r02_pytest_gpu_tail.txt
python tools/make_traffic_json.py $p/r02_particle_kernels_ncu.txt $p/r02_ncu_traffic_512.json
python - <<PY
import csv, collections
lines = [l for l in open("$p/r02_launches_bench512.csv") if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e3 if u in ("us", "usecond") else v / 1e6 if u in ("ns", "nsecond") else v
    agg.setdefault(r["Kernel Name"].split("(")[0][:90], []).append(v)
tot = sum(sum(v) for v in agg.values())
with open("$p/r02_launches_step_summary.txt", "w") as f:
    f.write("ncu launch list of bench.py --steps 2 (timed steps only, --profile-from-start off): ms per kernel, share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{sum(v):9.3f} ms {100 * sum(v) / tot:5.1f} %  n={len(v):3d}  {k}\n")
print(open("$p/r02_launches_step_summary.txt").read())
PY
