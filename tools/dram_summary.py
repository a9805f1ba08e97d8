"""Digest of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list (one or more generator forwards of `bench.py --no-graph`) into
profiles/r2_dram_per_launch.json: DRAM bytes and time per launch for every kernel class, next to
the algorithmic bytes of the styled-conv launches (planes in + planes out).

    python tools/dram_summary.py gpurun_out/launches.csv profiles/r2_dram_per_launch.json "<command>"
"""
import collections
import csv
import json
import re
import sys


def algorithmic_conv_bytes(batch=32):
    """bf16 hi+lo planes read + written by each styled conv of the 256^2 generator (4 B per
    element each way; layer 14 writes no planes, its ToRGB partials are 2 x 3 fp32 planes)."""
    chans = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128}
    out = {}
    cin, res, n = 512, 4, 2
    out['layer2'] = batch * 4 * (cin * 16 + 512 * 16)
    for r in (8, 16, 32, 64, 128, 256):
        cout = chans[r]
        out['layer%d' % (n + 1)] = batch * 4 * (cin * (r // 2) ** 2 + cout * r * r)
        out['layer%d' % (n + 2)] = batch * 4 * (cout * r * r + (cout * r * r if r < 256 else 0))
        n += 2
        cin = cout
    return out


def main(path, out_path, source):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        kid = row['ID']
        name = row['Kernel Name']
        m = re.search(r'rw::<unnamed>::(\w+)', name)
        name = ('rw::' + m.group(1)) if m else re.sub(r'[<(].*', '', name)[:60]
        ent = per.setdefault(kid, {'name': name})
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        if row['Metric Name'] == 'gpu__time_duration.sum':
            ent['ns'] = v * {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(unit, 1)
        elif row['Metric Name'].startswith('dram__bytes'):
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
            ent[row['Metric Name']] = v * mult
    classes = collections.OrderedDict()
    for ent in per.values():
        c = classes.setdefault(ent['name'], {'launches': 0, 'ns': 0.0, 'read': 0.0, 'write': 0.0})
        c['launches'] += 1
        c['ns'] += ent.get('ns', 0.0)
        c['read'] += ent.get('dram__bytes_read.sum', 0.0)
        c['write'] += ent.get('dram__bytes_write.sum', 0.0)
    conv = {'launches': 0, 'ns': 0.0, 'read': 0.0, 'write': 0.0}
    for k, c in classes.items():
        if 'conv_tc' in k or 'upconv_fused' in k:
            for f in conv:
                conv[f] += c[f]
    alg = sum(algorithmic_conv_bytes().values())
    forwards = max(1, round(conv['launches'] / 13.0))
    kernels = collections.OrderedDict()

    def pack(c, alg_bytes=None):
        d = {'launches': c['launches'], 'ms_total': c['ns'] / 1e6,
             'dram_bytes_per_launch': (c['read'] + c['write']) / max(1, c['launches']),
             'dram_read_bytes': c['read'], 'dram_write_bytes': c['write']}
        if alg_bytes is not None:
            d['algorithmic_bytes_per_launch'] = alg_bytes
            d['traffic_over_algorithmic'] = d['dram_bytes_per_launch'] / alg_bytes
        return d
    kernels['conv (conv_tc + upconv_fused)'] = pack(conv, alg * forwards / max(1, conv['launches']))
    for k, c in classes.items():
        kernels[k] = pack(c)
    res = {'source': source, 'forwards_captured': forwards, 'kernels': kernels,
           'note': 'ncu replays every kernel with cold caches; per-launch times are not bench '
                   'values, the DRAM byte counts are what is used'}
    with open(out_path, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(kernels['conv (conv_tc + upconv_fused)'], indent=1))
    for k, v in kernels.items():
        print('%-40s n=%3d  %8.3f ms  %8.1f MB/launch' % (k, v['launches'], v['ms_total'],
                                                          v['dram_bytes_per_launch'] / 1e6))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
