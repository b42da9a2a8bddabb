"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'cnld::<unnamed>::|void |cnld_capi::', '', name)[:48]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    print(f'# {path}: {sum(n for n, _ in agg.values())} launches, {tot:.1f} ms (cold-cache, serialised: compare shares)')
    print(f'{"kernel":50s} {"launches":>8s} {"ms":>10s} {"share":>7s}')
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{k:50s} {n:8d} {v:10.3f} {v / tot:7.3f}')


if __name__ == '__main__':
    main(sys.argv[1])
