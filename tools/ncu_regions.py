"""Source-level digest of one `ncu --set full --import-source on` capture (read here, on the
CPU container, from the .ncu-rep a gpurun visit brought back):

    python tools/ncu_regions.py gpurun_out/blur_l13.ncu-rep [--top 25]

Prints, for the first kernel of the report: duration / issue / DRAM headline numbers, the opcode
mix (share of executed warp instructions and of stall samples), the code regions (maximal SASS
runs with the same execution count = loop nests) with their share of the instruction stream, the
instructions with the most stall samples and their dominant stall reason, and the execution
counts of barrier polls / TMA / MMA instructions (how often each mbarrier wait spins).  This is
what told the blur kernel's index arithmetic (29 %), per-pixel tail (32 %) and tile decode
(11 %) apart, and showed the conv MMA issuer polling tmem_empty ~190 times per chunk.
"""
import collections
import csv
import io
import subprocess
import sys


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def headline(rep):
    keys = ['Duration', 'Elapsed Cycles', 'SM Frequency', 'DRAM Throughput', 'Memory Throughput',
            'Executed Ipc Active', 'Issue Slots Busy', 'Registers Per Thread', 'Achieved Occupancy',
            'L2 Hit Rate', 'highest-utilized']
    for ln in ncu(['-i', rep, '--page', 'details']).splitlines():
        if any(k in ln for k in keys):
            print('   ', ' '.join(ln.split()))


def source_rows(rep):
    txt = ncu(['-i', rep, '--page', 'source', '--csv'])
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else '?'
    hdr = rows[hdr_i]
    col = {n: hdr.index(n) for n in hdr}
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    out = []
    for r in rows[hdr_i + 1:]:
        try:
            out.append(dict(src=r[col['Source']].strip(), n=float(r[col['Instructions Executed']]),
                            s=float(r[col['# Samples']]),
                            stalls={c: float(r[col[c]] or 0) for c in stall_cols}))
        except (ValueError, IndexError):
            continue
    return kernel, out


def opcode(src):
    parts = src.split()
    if not parts:
        return '?'
    op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
    return op.split('.')[0]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 25
    kernel, rows = source_rows(rep)
    print('kernel:', kernel[:110])
    headline(rep)
    tot_n = sum(r['n'] for r in rows) or 1.0
    tot_s = sum(r['s'] for r in rows) or 1.0
    print('\nexecuted warp instructions: %.1f M over %d SASS instructions; %d stall samples'
          % (tot_n / 1e6, len(rows), tot_s))
    ops, ops_s = collections.Counter(), collections.Counter()
    for r in rows:
        ops[opcode(r['src'])] += r['n']
        ops_s[opcode(r['src'])] += r['s']
    print('\nopcode mix (executed %, stall samples %):')
    for op, n in ops.most_common(18):
        print('    %-10s %6.2f %%  %6.2f %%' % (op, 100 * n / tot_n, 100 * ops_s[op] / tot_s))
    print('\ncode regions (runs of equal execution count):')
    start, prev, acc = 0, None, 0.0
    regions = []
    for i, r in enumerate(rows):
        if prev is not None and abs(r['n'] - prev) > 0.02 * max(r['n'], prev, 1.0):
            regions.append((start, i - 1, prev, acc))
            start, acc = i, 0.0
        acc += r['n']
        prev = r['n']
    regions.append((start, len(rows) - 1, prev, acc))
    for a, b, n, acc in regions:
        if acc / tot_n >= 0.01:
            print('    sass %5d-%5d  x%-10.0f %5.1f %%   %s' % (a, b, n, 100 * acc / tot_n,
                                                              rows[a]['src'][:60]))
    print('\nmost stalled instructions:')
    for r in sorted(rows, key=lambda r: -r['s'])[:top]:
        reason = max(r['stalls'].items(), key=lambda kv: kv[1]) if r['stalls'] else ('', 0)
        print('    %5.1f %%  x%-10.0f %-62s %s' % (100 * r['s'] / tot_s, r['n'], r['src'][:62],
                                                  reason[0]))
    sync = [r for r in rows if any(t in r['src'] for t in ('SYNCS', 'UTCHMMA', 'UTCBAR', 'UTMALDG',
                                                          'LDTM', 'LDGSTS', 'BAR.SYNC'))]
    if sync:
        print('\nbarrier / TMA / MMA / async-copy instructions (execution counts):')
        seen = set()
        for r in sync:
            key = (r['src'][:70], r['n'])
            if key in seen or r['n'] == 0:
                continue
            seen.add(key)
            print('    x%-10.0f %s' % (r['n'], r['src'][:90]))


if __name__ == '__main__':
    main()
