"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys


def summarise(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name']
        m = re.search(r'rw::<unnamed>::(\w+)', name)
        name = ('rw::' + m.group(1)) if m else re.sub(r'[<(].*', '', name)[:60]
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(row['Metric Unit'], 1)
        tot[name] = tot.get(name, 0) + v
        cnt[name] += 1
    total = sum(tot.values())
    out = ['total captured %.3f ms, %d launches' % (total / 1e6, sum(cnt.values()))]
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:20]:
        out.append('%7.3f ms %5.1f%% n=%4d  %s' % (v / 1e6, 100 * v / total, cnt[k], k))
    return '\n'.join(out)


if __name__ == '__main__':
    print(summarise(sys.argv[1]))
