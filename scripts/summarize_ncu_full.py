"""Summarise an `ncu --set full` capture (exported with `ncu -i X.ncu-rep --page raw --csv`) per kernel:
launch count, duration, DRAM bytes, L2 / tensor-pipe / issue utilisation, registers, shared memory.

  python scripts/summarize_ncu_full.py gpurun_out/r01_tc_full_raw.csv profiles/r01_tc_full.md profiles/r01_ncu_full.json
"""
import collections
import csv
import json
import re
import sys


def unit_scale(u):
    u = u.split('/')[0]
    return {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1., 'ms': 1e3, 'usecond': 1., 'nsecond': 1e-3, 'msecond': 1e3}.get(u, 1.)


def main(src, md, js):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name, scaled=True):
        if name not in ix or r[ix[name]] in ('', 'n/a'):
            return None
        try:
            v = float(r[ix[name]].replace(',', ''))
        except ValueError:
            return None
        return v * unit_scale(units[ix[name]]) if scaled else v

    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '')
        key = (name, r[ix['Grid Size']], r[ix['Block Size']])
        a = agg.setdefault(key, collections.defaultdict(list))
        for k, m in (('us', 'gpu__time_duration.sum'), ('rd', 'dram__bytes_read.sum'), ('wr', 'dram__bytes_write.sum'),
                     ('l2', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), ('tensor', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
                     ('issue', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'), ('regs', 'launch__registers_per_thread'),
                     ('smem', 'launch__shared_mem_per_block_dynamic'), ('occ', 'sm__warps_active.avg.pct_of_peak_sustained_active')):
            full = m if m in ix else next((h for h in hdr if h.endswith(m)), None)
            v = col(r, full) if full else None
            if v is not None:
                a[k].append(v)
    mean = lambda v: sum(v) / len(v) if v else float('nan')
    out = {}
    with open(md, 'w') as f:
        f.write('| kernel | grid | block | launches | avg us | DRAM read / launch | DRAM write / launch | L2 % | tensor pipe % | issue % | warps active % | regs | dyn smem |\n')
        f.write('|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n')
        for (name, grid, block), a in agg.items():
            f.write('| `%s` | %s | %s | %d | %.1f | %.2f MB | %.3f MB | %.1f | %.2f | %.1f | %.1f | %d | %.0f KB |\n' % (
                name, grid, block, len(a['us']), mean(a['us']), mean(a['rd']) / 1e6, mean(a['wr']) / 1e6, mean(a['l2']), mean(a['tensor']),
                mean(a['issue']), mean(a['occ']), mean(a['regs']), mean(a['smem']) / 1e3))
            base = re.sub(r'<.*', '', name)
            o = out.setdefault(base, {'launches': 0, 'dram_bytes': 0., 'us': 0.})
            o['launches'] += len(a['us'])
            o['dram_bytes'] += sum(a['rd']) + sum(a['wr'])
            o['us'] += sum(a['us'])
    for o in out.values():
        o['dram_bytes_per_launch'] = o['dram_bytes'] / max(o['launches'], 1)
        o['avg_us'] = o['us'] / max(o['launches'], 1)
    json.dump(out, open(js, 'w'), indent=1)
    print(open(md).read())


if __name__ == '__main__':
    main(*sys.argv[1:4])
