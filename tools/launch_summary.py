#!/usr/bin/env python
"""Summarise an ncu launch-list CSV (gpu__time_duration.sum per launch) by kernel."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d['Metric Name'] == 'gpu__time_duration.sum':
                v = float(d['Metric Value'].replace(',', ''))
                u = d['Metric Unit']
                d['us'] = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
                data.append(d)
    return data


def short(name):
    return name.split('(')[0].replace('void ', '')[:70]


if __name__ == '__main__':
    data = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        k = short(d['Kernel Name'])
        agg[k][0] += 1
        agg[k][1] += d['us']
    tot = sum(v[1] for v in agg.values())
    print('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot * 100:.1f}% | {v[1] / v[0]:.1f} |')
    print(f'| total | {len(data)} | {tot:.1f} | 100% | |')
    if len(sys.argv) > 2:
        for d in data:
            if sys.argv[2] in d['Kernel Name']:
                print(f"{short(d['Kernel Name'])[:40]:42s} grid={d['Grid Size']:14s} {d['us']:8.1f}us")
