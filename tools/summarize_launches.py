#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table: per-kernel count, mean us, share."""
import collections, csv, re, sys


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        n = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '') + ' grid=' + d['Grid Size'].replace(' ', '')
        v = float(d['Metric Value'].replace(',', ''))
        u = d['Metric Unit']
        v = v / 1000. if u == 'ns' else (v * 1000. if u == 'ms' else v)
        agg.setdefault(n, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print('### %s\n' % title)
    print('%d launches, %.1f us total (cold-cache, serialised under ncu: compare SHARES, not absolutes)\n' % (len(data), tot))
    print('| kernel | launches | mean us | share |\n|---|---:|---:|---:|')
    for n, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        print('| `%s` | %d | %.1f | %.3f |' % (n, len(v), sum(v) / len(v), sum(v) / tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
