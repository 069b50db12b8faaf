"""SASS opcode histogram per kernel of libmcaz.so (cuobjdump -sass): the evidence that the tower is Blackwell-native --
UTCHMMA / UTCQMMA (tcgen05.mma bf16 / e4m3), LDTM (tcgen05.ld), UTMALDG (TMA loads), UTCBAR (tcgen05.commit), SYNCS (mbarrier) --
and what the other kernels are made of.  No GPU needed.   python tools/sass_histogram.py [libmcaz.so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'minitchess_alphazero_b200', 'libmcaz.so')
text = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
kernels = collections.OrderedDict()
name = None
for line in text.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace('(anonymous namespace)::', '').replace('mcaz::', '').replace('void ', '')
        name = re.sub(r'\(.*', '', name)
        kernels[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)', line)
    if m and name:
        op = m.group(1)
        base = op.split('.')[0]
        key = op if base.startswith(('UTC', 'UTMA', 'LDTM', 'STTM', 'UBLKCP', 'SYNCS', 'LDG', 'STG', 'HMMA', 'ATOM', 'RED', 'LDS', 'STS')) else base
        kernels[name][key] += 1
TENSOR = ('UTC', 'UTMA', 'LDTM', 'STTM', 'UBLKCP', 'SYNCS')
for name, c in kernels.items():
    total = sum(c.values())
    special = {k: v for k, v in c.items() if k.startswith(TENSOR)}
    print('%s: %d instructions' % (name, total))
    if special:
        print('   tcgen05 / TMA / mbarrier: ' + ', '.join('%s x%d' % kv for kv in sorted(special.items())))
    mem = {k: v for k, v in c.items() if k.startswith(('LDG', 'STG', 'ATOM', 'RED', 'LDS', 'STS'))}
    if mem:
        print('   memory: ' + ', '.join('%s x%d' % kv for kv in sorted(mem.items())))
    rest = [(k, v) for k, v in c.most_common(12) if not k.startswith(TENSOR)]
    print('   most frequent: ' + ', '.join('%s x%d' % kv for kv in rest))
