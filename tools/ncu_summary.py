#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + hottest source lines) without a GPU.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json out.json] [--top 40]
"""
import csv, io, json, subprocess, sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_elapsed.avg', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
STALLS = ['long_scoreboard', 'short_scoreboard', 'barrier', 'wait', 'branch_resolving', 'mio_throttle',
          'lg_throttle', 'not_selected', 'math_pipe_throttle', 'dispatch_stall', 'no_instruction', 'membar', 'sleeping']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        d = {k: (r[idx[k]] + ' ' + units[idx[k]]).strip() for k in KEEP if k in idx}
        for s in STALLS:
            k = f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio'
            if k in idx:
                d['stall_' + s] = r[idx[k]]
        out.append(d)
        for k, v in d.items():
            print(f'{k:85s} {v}')
        print('-' * 100)
    src = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv']))))
    hot = []
    if len(src) > 2:
        h = src[1]
        ix = {c: i for i, c in enumerate(h)}
        body = [r for r in src[2:] if r and r[0].startswith('0x') and len(r) >= len(h) - 2]
        seen, uniq = set(), []
        for r in body:
            if r[0] in seen:
                continue
            seen.add(r[0]); uniq.append(r)
        tot = sum(int(r[ix['# Samples']]) for r in uniq) or 1
        cols = [c for c in h if c.startswith('stall_') and '(' not in c]
        print(f'source: {len(uniq)} instructions, {tot} samples, '
              f'{sum(int(r[ix["Instructions Executed"]]) for r in uniq)} warp instructions executed')
        ranked = sorted(range(len(uniq)), key=lambda i: -int(uniq[i][ix['# Samples']]))[:top]
        for i in sorted(ranked):
            r = uniq[i]
            s = int(r[ix['# Samples']])
            tops = sorted(((int(r[ix[c]]), c[6:]) for c in cols), reverse=True)[:2]
            line = {'index': i, 'samples_pct': round(100 * s / tot, 2), 'sass': r[ix['Source']].strip(),
                    'executed': int(r[ix['Instructions Executed']]),
                    'top_stalls': ' '.join(f'{c}={v}' for v, c in tops if v)}
            hot.append(line)
            print(f"{i:5d} {line['samples_pct']:5.2f}% ex={line['executed']:9d} {line['sass'][:64]:64s} {line['top_stalls']}")
    if '--json' in sys.argv:
        json.dump({'report': rep, 'launches': out, 'hot_instructions': hot},
                  open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)


if __name__ == '__main__':
    main()
