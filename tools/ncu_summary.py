#!/usr/bin/env python
"""Condenses an .ncu-rep (read with `ncu -i`, no GPU needed) into the short text summary kept under profiles/:
key metrics per profiled launch, warp-stall breakdown, hottest stall sites and the dynamic instruction mix.

  python tools/ncu_summary.py gpurun_out/prof_pairs.ncu-rep > profiles/r01_pairs_full.txt
  python tools/ncu_summary.py --launches gpurun_out/launches.csv > profiles/r01_launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size',
    'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sectors_srcunit_tex_op_red.sum', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_report(rep):
    rows = ncu_csv(rep, 'raw')
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {n: i for i, n in enumerate(hdr)}
    print(f'# {rep}: {len(data)} profiled launch(es)')
    for r in data:
        print(f"\n## {r[ci['Kernel Name']][:150]}")
        for k in KEYS:
            if k in ci:
                print(f'{k:72s} {r[ci[k]]:>18s} {units[ci[k]]}')
        traffic = None
        try:
            rd, wr = float(r[ci['dram__bytes_read.sum']]), float(r[ci['dram__bytes_write.sum']])
            mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
            traffic = rd * mult[units[ci['dram__bytes_read.sum']]] + wr * mult[units[ci['dram__bytes_write.sum']]]
            print(f"{'dram traffic (read+write) per launch':72s} {traffic / 1e9:18.4f} GB")
        except (KeyError, ValueError):
            pass
        print('-- warps stalled per issue-active cycle (smsp__average_warps_issue_stalled_*_per_issue_active)')
        st = [(h.split('stalled_')[1].split('_per_issue')[0], float(r[i])) for h, i in ci.items()
              if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and r[i]]
        for name, v in sorted(st, key=lambda kv: -kv[1])[:9]:
            print(f'   {name:28s} {v:6.2f}')
    src = ncu_csv(rep, 'source', ('--print-source', 'sass'))
    hidx = [i for i, r in enumerate(src) if r and r[0] == 'Address']
    if not hidx:
        return
    h = src[hidx[0]]
    body = src[hidx[0] + 1:(hidx[1] - 1 if len(hidx) > 1 else len(src))]
    c = {n: i for i, n in enumerate(h)}
    total = sum(int(r[c['# Samples']] or 0) for r in body)
    print(f'\n## SASS-level view of the first launch: {len(body)} static instructions, {total} stall samples')
    for col in ('stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_short_sb'):
        if col not in c:
            continue
        tot = sum(int(r[c[col]] or 0) for r in body)
        print(f'-- {col}: {tot} samples ({100.0 * tot / max(total, 1):.1f}%); hottest sites:')
        for r in sorted(body, key=lambda r: -int(r[c[col]] or 0))[:4]:
            print(f"   {r[c[col]]:>7s}  {r[c['Source']][:90]}")
    mix = collections.Counter()
    for r in body:
        toks = r[c['Source']].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        mix[op.split('.')[0]] += int(r[c['Instructions Executed']] or 0)
    t = sum(mix.values())
    print(f'-- dynamic warp-instruction mix ({t} warp instructions)')
    print('   ' + '  '.join(f'{k} {100.0 * v / t:.1f}%' for k, v in mix.most_common(16)))


def summarize_launches(path):
    rows = list(csv.reader(open(path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        try:
            v = float(r[-1].replace(',', ''))
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(r[4][:110], [0, 0.0])
        a[0] += 1
        a[1] += v
    unit = rows[start][-1] if rows[start] else ''
    tot = sum(t for _, t in agg.values())
    print(f'# {path}: per-kernel totals over the profiled run (gpu__time_duration.sum, ns; cold-cache, serialised)')
    print(f'{"total_ms":>10s} {"share":>7s} {"count":>6s} {"avg_us":>10s}  kernel')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{t / 1e6:10.3f} {100 * t / tot:6.1f}% {n:6d} {t / n / 1e3:10.1f}  {k}')


if __name__ == '__main__':
    if sys.argv[1] == '--launches':
        summarize_launches(sys.argv[2])
    else:
        summarize_report(sys.argv[1])
