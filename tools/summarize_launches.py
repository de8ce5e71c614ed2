"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

usage: python tools/summarize_launches.py launches.csv [skip_first_n_launches [count]]
"""
import csv, collections, re, sys

def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    count = int(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    h = rows[hdr]
    kn, mv, gs = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[hdr + 1 + skip:][:count]:
        if len(r) <= mv:
            continue
        name = re.sub(r'\(.*$', '', r[kn])
        name = re.sub(r'^void ', '', name)
        t = float(r[mv].replace(',', '')) / 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    print(f'{"kernel":90s} {"n":>5s} {"ms":>9s} {"share":>6s}')
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{name[:90]:90s} {n:5d} {t:9.3f} {100*t/total:5.1f}%')
    print(f'{"TOTAL":90s} {sum(a[0] for a in agg.values()):5d} {total:9.3f}')

if __name__ == '__main__':
    main()
