#!/usr/bin/env python
"""Where a kernel's issued instructions and warp samples go, from `ncu -i X.ncu-rep --page source --csv --print-source sass
--launch-skip K --launch-count 1 > src.csv`: consecutive SASS instructions with the same executed count form a region (a loop level);
prints per region its address range, executions per instruction, share of issued instructions and of warp samples, and the top
stall reasons.    python scripts/ncu_regions.py src.csv [units]   (units: e.g. the number of targets, for per-unit figures)"""
import csv
import sys


def main(path, units=None):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if 'Address' in r and 'Source' in r)
    ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stalls = [(h[6:], hdr.index(h)) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows:
        if len(r) >= len(hdr) and r[ia].startswith('0x'):
            data.append((int(r[ia], 16), r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0), [int(r[i] or 0) for _, i in stalls]))
    seen, uniq = set(), []
    for d in data:                                   # the page lists the kernel once per launch selected; keep the first copy
        if d[0] in seen:
            break
        seen.add(d[0]); uniq.append(d)
    data = uniq
    base = data[0][0]
    tot = sum(d[2] for d in data) or 1
    ts = sum(d[3] for d in data) or 1
    print(f'# {path}: {len(data)} SASS instructions, {tot / 1e6:.2f} M issued, {ts} warp samples' + (f', {tot / units:.1f} issued per unit' if units else ''))
    print('# region (byte offsets)   exec/instr     issued M    %   samples    %   per unit   top stalls')
    start, prev, ex_acc, sm_acc, st_acc = 0, None, 0, 0, [0] * len(stalls)

    def flush(end):
        if ex_acc < 0.004 * tot and sm_acc < 0.004 * ts:
            return
        top = sorted(zip([n for n, _ in stalls], st_acc), key=lambda kv: -kv[1])[:3]
        tsum = sum(st_acc) or 1
        print(f'{start:6x}-{end:6x} {prev / 1e6:10.3f}M {ex_acc / 1e6:10.2f} {100 * ex_acc / tot:5.1f} {sm_acc:8d} {100 * sm_acc / ts:5.1f} '
              f'{(ex_acc / units if units else 0):9.1f}   ' + ', '.join(f'{n} {100 * v / tsum:.0f}%' for n, v in top if v))

    for a, s, ex, sm, st in data:
        off = a - base
        if prev is not None and abs(ex - prev) > 0.05 * max(ex, prev, 1):
            flush(off)
            start, ex_acc, sm_acc, st_acc = off, 0, 0, [0] * len(stalls)
        ex_acc += ex; sm_acc += sm
        st_acc = [x + y for x, y in zip(st_acc, st)]
        prev = ex
    flush(data[-1][0] - base + 16)


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
