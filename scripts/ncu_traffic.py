#!/usr/bin/env python
"""profiles/gather_traffic.json from an `ncu --set full` capture of the gather launches of one bench step (4 since round 2: one per cell):
    ncu -i gpurun_out/prof_gather_step.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_traffic.py /tmp/raw.csv N_GRAINS > profiles/gather_traffic.json"""
import csv
import json
import sys


def main(path, n_grains):
    rows = list(csv.reader(open(path)))
    head, data = rows[0], rows[2:]
    units = rows[1]
    ir, iw, it, ik = (head.index(k) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'Kernel Name'))
    scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
    tot = sum(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in data)
    print(json.dumps({'n_grains': n_grains, 'launches': len(data), 'dram_bytes_per_launch': tot / len(data),
                      'dram_bytes_per_step': tot, 'ncu_us_per_launch': sum(float(r[it]) for r in data) / len(data),
                      'kernels': sorted({r[ik].split('(')[0] for r in data}),
                      'source': 'ncu --set full --clock-control none, one step of bench.py (cold-cache, serialised launches)'}))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]))
