"""Kernel timeline of graph-replayed generator steps (torch.profiler / CUPTI): busy time per kernel
name and the idle gaps between kernels inside a step — what ncu's serialised cold-cache launch list
cannot show.

    python tools/prof_step_timeline.py [steps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import collections   # noqa: E402
import torch         # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    import bench
    from rewriting_b200.graphs import GraphedModule
    dev = torch.device('cuda')
    model = bench.build_model(dev)
    z = torch.randn(bench.BATCH, 512, device=dev)
    runner = GraphedModule(model, z)
    for _ in range(5):
        runner(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        runner(z)
    e1.record()
    torch.cuda.synchronize()
    print('unprofiled: %.3f ms per step' % (e0.elapsed_time(e1) / 20))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            runner(z)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.cuda_time_total >= 0]
    ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in evs if 'emcpy' not in e.name and 'emset' not in e.name])
    if not ks:
        print('no CUDA kernel events (CUPTI unavailable?)')
        return
    span = (ks[-1][1] - ks[0][0]) / 1e3
    busy = collections.Counter()
    cnt = collections.Counter()
    gaps = []
    for i, (s, e, n) in enumerate(ks):
        key = n.split('(')[0].split('<')[0][-48:]
        busy[key] += (e - s) / 1e3
        cnt[key] += 1
        if i:
            gaps.append((max(0.0, s - ks[i - 1][1]) / 1e3, ks[i - 1][2].split('(')[0][-30:], key))
    tot = sum(busy.values())
    print('%d kernels over %d steps: span %.3f ms (%.3f per step), busy %.3f ms, idle %.3f ms (%.1f %%)' % (
        len(ks), steps, span, span / steps, tot, span - tot, 100 * (span - tot) / span))
    for k, v in busy.most_common(12):
        print('  %9.3f ms/step  n/step %5.1f  %s' % (v / steps, cnt[k] / steps, k))
    g = sorted(gaps, reverse=True)[:8]
    print('largest gaps (ms): ' + ', '.join('%.4f after %s' % (a, b) for a, b, _ in g))
    import statistics
    print('median gap %.2f us, mean %.2f us' % (1e3 * statistics.median(x[0] for x in gaps), 1e3 * sum(x[0] for x in gaps) / len(gaps)))


if __name__ == '__main__':
    main()
