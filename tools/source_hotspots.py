"""Instructions executed and stall samples per source function / line from an ncu report captured with --import-source on
(`ncu --set full --import-source on -k regex:<kernel> -o rep ...`); needs ncu on PATH, no GPU.
    python tools/source_hotspots.py gpurun_out/x.ncu-rep [top_lines] > profiles/rNN_x_hotspots.txt"""
import collections
import csv
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


cur, hdr, idx = None, None, {}
inst, samples, src = collections.Counter(), collections.Counter(), {}
for r in rows:
    if r and r[0] == 'File Path':
        cur = os.path.basename(r[1])
    elif r and r[0] == 'Line No':
        hdr, idx = r, {}
        for i, k in enumerate(r):
            idx.setdefault(k, i)
    elif hdr and len(r) >= len(hdr) and r[0].isdigit():
        key = (cur, int(r[0]))
        inst[key] += num(r[idx['Instructions Executed']])
        samples[key] += num(r[idx['# Samples']])
        src[key] = r[1].strip()[:100]
ti, ts = sum(inst.values()), sum(samples.values())
print('%s: %d warp instructions executed, %d stall samples (all captured launches)' % (os.path.basename(rep), ti, ts))
print('\nby function (definitions found in minitchess_alphazero_b200/csrc):')
by_fn = collections.Counter()
for fn in sorted({f for f, _ in inst}):
    path = os.path.join(REPO, 'minitchess_alphazero_b200', 'csrc', fn)
    if not os.path.exists(path):
        by_fn[(fn, 0, '(header)')] += sum(n for (f, _), n in inst.items() if f == fn)
        continue
    starts = []
    for i, line in enumerate(open(path).read().split('\n'), 1):
        if re.match(r'^(MC_HD|__device__|__global__|template|static|inline)\b', line) and '(' in line and not line.rstrip().endswith(';'):
            name = re.findall(r'([A-Za-z_0-9]+)\s*\(', line)
            starts.append((i, name[0] if name else line[:30]))
    starts.append((10 ** 9, 'END'))
    for (f, l), n in inst.items():
        if f != fn:
            continue
        for k in range(len(starts) - 1):
            if starts[k][0] <= l < starts[k + 1][0]:
                by_fn[(fn, starts[k][0], starts[k][1])] += n
                break
        else:
            by_fn[(fn, 0, '(before the first function)')] += n
for (fn, l, name), n in by_fn.most_common(28):
    print('   %-18s %5d %-30s %9d %5.1f %%' % (fn, l, name, n, 100.0 * n / max(ti, 1)))
print('\nby line, instructions executed:')
for (f, l), n in inst.most_common(top):
    print('   %-18s %5d %9d %5.1f %%  samples %5d | %s' % (f, l, n, 100.0 * n / max(ti, 1), samples[(f, l)], src[(f, l)]))
print('\nby line, stall samples:')
for (f, l), n in samples.most_common(top):
    print('   %-18s %5d %6d %5.1f %% | %s' % (f, l, n, 100.0 * n / max(ts, 1), src[(f, l)]))
