"""Per-kernel summary of one HFRM engine call from an ncu CSV (gpu__time_duration + dram bytes) of tools/bench_hfrm.py."""
import collections
import csv
import re
import sys


def main(path, detail=False):
    rows = [l for l in open(path) if not l.startswith('==')]
    by = collections.OrderedDict()
    for x in csv.DictReader(rows):
        d = by.setdefault(int(x['ID']), {'name': x['Kernel Name'], 'grid': x['Grid Size']})
        d[x['Metric Name']] = float(x['Metric Value'].replace(',', '')) * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}.get(x['Metric Unit'].lower(), 1)
    ids = list(by)
    ci = [k for k, i in enumerate(ids) if 'conv_in' in by[i]['name']]
    co = [k for k, i in enumerate(ids) if 'conv_out' in by[i]['name']]
    sel = [by[i] for i in ids[ci[-1]:co[-1] + 1]] if co and co[-1] > ci[-1] else [by[i] for i in ids[ci[-2]:ci[-1]]]
    agg = collections.OrderedDict()
    for n, d in enumerate(sel):
        nm = re.sub(r'.*(hfrm_|gemm_)', r'\1', d['name'])
        nm = re.sub(r'_kernel.*', '', nm)
        t = d['gpu__time_duration.sum'] / 1e3
        bt = d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']
        if detail:
            print(f"{n:3d} {nm:14s} {d['grid']:>16s} {t:8.1f} us {bt / 1e6:8.1f} MB {bt / t / 1e3:7.0f} GB/s")
        a = agg.setdefault((nm, d['grid']), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += bt
    tot = sum(a[1] for a in agg.values())
    print(f"launches {len(sel)}  total {tot:.1f} us (cold-cache, serialised)")
    for k, (c, t, bt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us {100 * t / tot:5.1f}% n={c:3d} {t / c:8.1f} us/launch {bt / c / 1e6:8.1f} MB/launch {bt / t / 1e3:7.0f} GB/s  {k[0]} {k[1]}")


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
