#!/usr/bin/env python
"""profiles/r02_ncu_traffic_512.json from an ncu summary (tools/ncu_summary.py output of the `--set full` capture of one
512^3 step): DRAM bytes (read + write) per C-ABI call = the sum over the kernels that call launches, each averaged
over its captured launches.   usage: python tools/make_traffic_json.py summary.txt out.json"""
import json
import re
import sys

KERNEL_TO_CALL = {
    "step_sort_kernel": "psc_step_sort", "step_sort_local_kernel": "psc_step_sort",
    "deposit_binned_kernel": "psc_deposit_sorted", "interp_kick_phi_binned_kernel": "psc_interp_kick_phi_sorted",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
per_kernel = {}
name = None
for line in open(sys.argv[1]):
    m = re.match(r"==\s+(?:void\s+)?(\w+)", line)
    if m:
        name = m.group(1)
        per_kernel.setdefault(name, []).append(0.0)
        continue
    m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([\d.]+)\s+(\w+)", line)
    if m and name:
        per_kernel[name][-1] += float(m.group(2)) * UNIT[m.group(3)]
out = {}
for k, v in per_kernel.items():
    call = KERNEL_TO_CALL.get(k)
    if call:
        out[call] = out.get(call, 0.0) + sum(v) / len(v)
out["source"] = ("ncu --set full --clock-control none, one 512^3 step (tools/prof_step.py), dram__bytes_read.sum + "
                 "dram__bytes_write.sum per launch, summed over the kernels of the call: " + sys.argv[1])
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
