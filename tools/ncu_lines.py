"""Per-source-line digest of an `ncu --set full --import-source on` capture (kernels built with
-lineinfo): executed warp instructions and stall samples attributed to each CUDA source line.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [--top 40] [--file upconv_tc.cu]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    only = sys.argv[sys.argv.index('--file') + 1] if '--file' in sys.argv else None
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    inst = collections.Counter()
    stall = collections.Counter()
    src = {}
    cur_file = None
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] in ('File Path', 'File Name'):
            cur_file = r[1].split('/')[-1]
            hdr = None
            continue
        if r and r[0] == 'Line No':
            hdr = {}
            for i, n in enumerate(r):
                hdr.setdefault(n, i)          # first 'Source' column = the CUDA line
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            line = int(r[hdr['Line No']])
            n = int(float(r[hdr['Instructions Executed']] or 0))
            s = int(float(r[hdr['# Samples']] or 0))
        except ValueError:
            continue
        key = (cur_file, line)
        inst[key] += n
        stall[key] += s
        src[key] = r[hdr['Source']].strip()[:110]
    tot_i = sum(inst.values()) or 1
    tot_s = sum(stall.values()) or 1
    print('total %.1f M warp instructions, %d stall samples' % (tot_i / 1e6, tot_s))
    keys = [k for k in inst if only is None or k[0] == only]
    for k in sorted(keys, key=lambda k: -inst[k])[:top]:
        print('%5.1f%% inst %5.1f%% stall  %s:%d  %s' % (100.0 * inst[k] / tot_i, 100.0 * stall[k] / tot_s,
                                                        k[0], k[1], src[k]))


if __name__ == '__main__':
    main()
