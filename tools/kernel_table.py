"""Per-kernel roofline table (markdown) from committed ncu raw pages: time, DRAM traffic and GB/s against the measured HBM
peak, threads per warp instruction, warp slots, registers; tensor-pipe utilisation for the tower.
    python tools/kernel_table.py profiles/r02_hot_raw.csv profiles/r02_all_raw.csv ... > profiles/r02_kernel_table.md
The launch with the median duration stands for each kernel of a page."""
import csv
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}


def peak_hbm():
    try:
        d = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        for k in ('hbm_gbs', 'hbm_gb_s', 'hbm_copy_gbs'):
            if k in d:
                return float(d[k])
        for v in d.values():
            if isinstance(v, dict):
                for k, x in v.items():
                    if 'hbm' in k.lower() and isinstance(x, (int, float)):
                        return float(x)
    except Exception:
        pass
    return 6547.5


def short(name):
    name = re.sub(r'\(.*', '', name)
    name = name.replace('void ', '').replace('unnamed>::', '').replace('<unnamed>::', '')
    return name.split('::')[-1].strip()


def rows_of(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, key, scale=True):
        if key not in idx or r[idx[key]] in ('', 'n/a'):
            return None
        v = float(r[idx[key]].replace(',', ''))
        return v * UNIT.get(units[idx[key]], 1.0) if scale else v
    out = {}
    for r in rows[2:]:
        k = short(r[idx['Kernel Name']])
        out.setdefault(k, []).append({
            't': val(r, 'gpu__time_duration.sum'), 'rd': val(r, 'dram__bytes_read.sum'), 'wr': val(r, 'dram__bytes_write.sum'),
            'tpi': val(r, 'smsp__thread_inst_executed_per_inst_executed.ratio', False),
            'warps': val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active', False),
            'regs': val(r, 'launch__registers_per_thread', False),
            'tensor': val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', False),
            'l2hit': val(r, 'lts__t_sector_hit_rate.pct', False),
            'inst': val(r, 'smsp__inst_executed.sum', False),
            'grid': r[idx['Grid Size']].replace(' ', ''), 'block': r[idx['Block Size']].replace(' ', '')})
    return out


def main():
    hbm = peak_hbm()
    print('| Kernel | page | launch | time | DRAM read + write | GB/s (%% of %.0f) | threads / warp instr. | warp slots | regs | L2 hit | tensor pipe |' % hbm)
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    for path in sys.argv[1:]:
        for k, ls in sorted(rows_of(path).items()):
            ls.sort(key=lambda x: x['t'])
            m = ls[len(ls) // 2]
            gbs = (m['rd'] + m['wr']) / m['t'] / 1e9
            print('| `%s` | %s | %s x %s | %.1f us | %.2f + %.2f MB | %.0f (%.1f %%) | %.1f / 32 | %.0f %% | %d | %.0f %% | %s |' % (
                k, os.path.basename(path).replace('_raw.csv', ''), m['grid'], m['block'], m['t'] * 1e6, m['rd'] / 1e6, m['wr'] / 1e6, gbs,
                100 * gbs / hbm, m['tpi'], m['warps'], m['regs'], m['l2hit'] or 0, ('%.1f %%' % m['tensor']) if m['tensor'] else '-'))


if __name__ == '__main__':
    main()
