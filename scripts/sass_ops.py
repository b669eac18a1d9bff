#!/usr/bin/env python
"""Count the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md, "What proves a Blackwell-native kernel")
per object file of the in-tree library and per kernel:
    python scripts/sass_ops.py > profiles/sass_ops.txt
UTCHMMA = tcgen05.mma (kind::tf32), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,
SYNCS = mbarrier, FFMA2 = packed fp32 FMA (sm_100), HMMA = legacy mma.sync (none expected)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'graingraphnn_b200', 'lib')
OPS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'FFMA2', 'FMUL2', 'HMMA', 'USETMAXREG', 'ATOM', 'RED']


def main():
    objs = sorted(f for f in os.listdir(LIB) if f.endswith('.o'))
    print('# cuobjdump -sass of graingraphnn_b200/lib/*.o (sm_100a), mnemonic counts per kernel; lines with no listed op are omitted')
    print('# ' + ' '.join(OPS))
    total = collections.Counter()
    for o in objs:
        out = subprocess.run(['cuobjdump', '-sass', os.path.join(LIB, o)], capture_output=True, text=True).stdout
        arch = re.findall(r'arch = (sm_\w+)', out)
        kern, per = None, collections.OrderedDict()
        for line in out.splitlines():
            m = re.match(r'\s*Function : (\S+)', line)
            if m:
                kern = m.group(1)
                per[kern] = collections.Counter()
                continue
            if kern is None:
                continue
            m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
            if m:
                op = m.group(1)
                for name in OPS:
                    if op == name or op.startswith(name + '.') or (name in ('ATOM', 'RED') and op.startswith(name)):
                        per[kern][name] += 1
        print(f'\n== {o}  ({", ".join(sorted(set(arch)))})')
        for k, c in per.items():
            if c:
                name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip() or k
                name = re.sub(r'\(anonymous namespace\)::', '', name)
                print(f'  {name[:110]:110s} ' + ' '.join(f'{n}={c[n]}' for n in OPS if c[n]))
                total.update(c)
    print('\n== library total: ' + ' '.join(f'{n}={total[n]}' for n in OPS))
    return 0


if __name__ == '__main__':
    sys.exit(main())
