"""Summarise Nsight Compute CSV exports for profiles/:

  python tools/ncu_summary.py launches gpurun_out/r02a_launches_step_t6.csv      # share of one step by kernel
  python tools/ncu_summary.py full gpurun_out/r02a_ncu_full_vit_t6.csv [...]     # one row per kernel: time, DRAM bytes, pipes

`launches` reads the `--metrics gpu__time_duration.sum --csv` log, `full` the `--page raw --csv` export of a --set full
report.  Kernel names are shortened to the function name + template arguments."""
import csv
import io
import json
import re
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'^void ', '', name)
    m = re.match(r'([\w:]+(<[^()]*>)?)', name)
    return m.group(1) if m else name


def read_csv(path):
    lines = open(path, errors='replace').read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.reader(io.StringIO('\n'.join(lines[start:]))))


def launches(path):
    rows = read_csv(path)
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        v = v / 1e3 if r[iu] in ('ns', 'nsecond') else (v * 1e3 if r[iu] in ('ms', 'msecond') else v)
        a = agg[short(r[ik])]
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f'| kernel | launches / step | us / step | share |\n|---|---|---|---|')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {n} | {us:.0f} | {us / tot * 100:.1f} % |')
    print(f'| total | {sum(a[0] for a in agg.values())} | {tot:.0f} | |')


WANT = OrderedDict([
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active', 'hmma%'),
    ('sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'umma%'),
    ('sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'umma2%'),
    ('sm__inst_executed.avg.per_cycle_elapsed', 'ipc'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu_wave%'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_waves'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
])


def full(paths):
    out_rows, traffic = [], {}
    for path in paths:
        rows = read_csv(path)
        hdr, units = rows[0], rows[1]
        ik = hdr.index('Kernel Name')
        cols = {m: hdr.index(m) for m in WANT if m in hdr}
        extra = [h for h in hdr if 'umma' in h.lower() or 'tensor' in h.lower()]
        for r in rows[2:]:
            if len(r) <= ik:
                continue
            d = OrderedDict(kernel=short(r[ik]))
            for m, i in cols.items():
                try:
                    d[WANT[m]] = float(r[i].replace(',', ''))
                    d[WANT[m] + '_unit'] = units[i]
                except ValueError:
                    pass
            out_rows.append(d)
    if not out_rows:
        return
    print('| kernel | grid | regs | time (us) | DRAM rd+wr (MB) | dram % | tensor % | ipc | lsu wavefront % | occupancy % |')
    print('|---|---|---|---|---|---|---|---|---|---|')
    for d in out_rows:
        t = d.get('time', 0.0)
        t_us = t / 1e3 if d.get('time_unit', 'ns').startswith('n') else (t * 1e3 if d.get('time_unit', '').startswith('m') else t)
        def mb(key):
            v, u = d.get(key, 0.0), d.get(key + '_unit', 'byte')
            return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)
        tens = max(d.get('umma%', 0.0), d.get('umma2%', 0.0), d.get('hmma%', 0.0))
        print(f"| `{d['kernel']}` | {int(d.get('grid', 0))} | {int(d.get('regs', 0))} | {t_us:.1f} | {mb('dram_rd') + mb('dram_wr'):.1f} | "
              f"{d.get('dram%', 0):.1f} | {tens:.1f} | {d.get('ipc', 0):.2f} | {d.get('lsu_wave%', 0):.1f} | {d.get('occ%', 0):.1f} |")
        traffic.setdefault(d['kernel'], []).append(dict(us=round(t_us, 2), dram_mb=round(mb('dram_rd') + mb('dram_wr'), 2)))
    json.dump(traffic, open('/tmp/ncu_traffic.json', 'w'), indent=1)


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
