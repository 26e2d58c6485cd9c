"""SASS opcode evidence per kernel of libibinn_b200.so (what proves the Blackwell-native paths, B200_PROFILING.md): counts of
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP / UTMALDG (TMA bulk copies), SYNCS (mbarrier), HMMA (legacy mma.sync),
FFMA, and the instruction total, per kernel entry point.

    python scripts/sass_histogram.py [ibinn_b200/csrc/libibinn_b200.so] > profiles/r02_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'ibinn_b200', 'csrc', 'libibinn_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
pat = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)')
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name.replace('(anonymous namespace)::', '').replace('void ', ''))
        kern = name
        counts[kern] = collections.Counter()
        continue
    m = pat.match(line)
    if m and kern:
        op = m.group(1).split('.')[0]
        counts[kern]['total'] += 1
        for key in ('UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'SYNCS', 'HMMA', 'FFMA', 'FFMA2', 'MUFU', 'ATOM', 'RED'):
            if op == key or (key == 'ATOM' and op.startswith('ATOM')):
                counts[kern][key] += 1
cols = ['UTCHMMA', 'LDTM', 'UBLKCP', 'SYNCS', 'HMMA', 'FFMA', 'FFMA2', 'MUFU', 'ATOM', 'RED', 'total']
print('# SASS opcode histogram of `%s` (sm_100a)\n' % os.path.relpath(lib, ROOT))
print('`UTCHMMA` = tcgen05.mma (kind::tf32), `LDTM` = tcgen05.ld, `UBLKCP` = cp.async.bulk (TMA), `SYNCS` = mbarrier ops; `HMMA` would be the legacy '
      'mma.sync path (absent).\n')
print('| kernel | ' + ' | '.join(cols) + ' |')
print('|---|' + '---:|' * len(cols))
for k, c in sorted(counts.items(), key=lambda kv: (-kv[1]['UTCHMMA'], -kv[1]['total'])):
    print('| `%s` | ' % k[:70] + ' | '.join(str(c[x]) for x in cols) + ' |')
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print('\nlibrary totals: ' + ', '.join('%s %d' % (x, tot[x]) for x in cols))
