"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/profile_unet.py: per-kernel totals
of the LAST UNet forward in the log (starts at the last timestep-embedding `linear_kernel` triple)."""
import collections
import csv
import re
import sys


def main(path, per_launch=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [x['Kernel Name'] for x in rows]
    idx = [i for i, n in enumerate(names) if 'linear_kernel' in n]
    sel = rows[idx[-3]:] if len(idx) >= 3 else rows
    agg = collections.OrderedDict()
    tot = 0.0
    for x in sel:
        n = re.sub(r'\(.*', '', x['Kernel Name'])
        n = re.sub(r'^void ', '', n)
        n = n.replace('wdm::(anonymous namespace)::', '').replace('wdm::<unnamed>::', '')
        t = float(x['Metric Value']) / 1e3
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
        if per_launch and ('gemm' in n):
            print(f"{t:8.1f} us grid={x['Grid Size']:>14s} {n[:70]}")
    print(f"launches in one UNet forward: {len(sel)}   total {tot:.1f} us (cold-cache, serialised: compare shares)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us {100 * t / tot:5.1f}%  n={c:3d}  {n[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
