#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel and shared-memory wavefronts / stalls per SASS opcode.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else None
WANT = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'launch__block_size']


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
for r in raw[2:]:
    name = r[hdr.index('Kernel Name')]
    if rx and rx not in name:
        continue
    print('\n== ', name[:110])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'   {w:70s} {r[i]:>20s} {units[i]}')
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--kernel-name", "regex:" + name.split('(')[0].split('<')[0].split()[-1]]))))
    if len(src) < 3:
        continue
    h = src[1]
    try:
        iS, iW, iI, iE = h.index('Source'), h.index('L1 Wavefronts Shared'), h.index('L1 Wavefronts Shared Ideal'), h.index('Instructions Executed')
    except ValueError:
        continue
    agg = {}
    for row in src[2:]:
        try:
            w, i, e = int(row[iW]), int(row[iI]), int(row[iE])
        except (ValueError, IndexError):
            continue
        if w == 0:
            continue
        toks = row[iS].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        a = agg.setdefault(op, [0, 0, 0])
        a[0] += w; a[1] += i; a[2] += e
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'   shared {k:10s} wavefronts {v[0]:>12d} ideal {v[1]:>12d} instr {v[2]:>11d}  wf/instr {v[0] / max(v[2], 1):5.2f} (ideal {v[1] / max(v[2], 1):4.2f})')
    stall = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
    s = {x: 0 for x in stall}
    for row in src[2:]:
        for x in stall:
            try:
                s[x] += int(row[h.index(x)])
            except (ValueError, IndexError):
                pass
    tot = sum(s.values()) or 1
    print('   stalls: ' + ', '.join(f'{k[6:]} {100 * v / tot:.0f}%' for k, v in sorted(s.items(), key=lambda kv: -kv[1])[:7]))
