#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + the most-sampled SASS instructions with their dominant stall reasons.
usage: python tools/ncu_top.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[-1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_xu.sum', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pcsamp_sample_buffer_full']
for i, n in enumerate(h):
    if n in want:
        print(f"{n:75s} {v[i]:>16s} {u[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, x in enumerate(rows) if x and x[0] == "Address")
h = rows[hi]; data = [x for x in rows[hi + 1:] if len(x) == len(h)]
isrc = h.index('Source'); isamp = h.index('# Samples'); iex = h.index('Instructions Executed')
stall = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(x[isamp] or 0) for x in data)
print('total samples', tot)
agg = {}
for x in data:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(x[i] or 0)
print('stall totals', dict(sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for x in sorted(data, key=lambda x: -int(x[isamp] or 0))[:N]:
    s = {h[i]: int(x[i] or 0) for i in stall if int(x[i] or 0) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(x[isamp].rjust(7), x[iex].rjust(10), x[isrc][:80].ljust(80), s)
