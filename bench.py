#!/usr/bin/env python
"""bench.py — headline benchmark of the rewriting_b200 hot path (driver contract).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

metric   : StyleGAN2-256 images/sec (BASELINE.json), synthetic random z, seeded random weights
workload : SeqStyleGAN2(256, mconv='seq') generator forward, batch=32 per GPU, fp32 in/out,
           conv operands 3-term split bf16 on tcgen05 tensor cores (fp32 accumulate)
step     : one batch of 32 latents through the whole generator -> 32 images (per GPU)
value    : images/s with z resident in HBM, CUDA-event timed, max over ranks, whole job
e2e      : same through the public API with HOST buffers: pinned z -> H2D, model(z), D2H of
           the images into pinned memory, inside the timed region
extra    : key-covariance samples/sec (layer 8, BASELINE.json's second metric): context
           forward + tensor-core second moment per batch of z, one all-reduce of (mom2,count)
           at the end; rewrite-loop iterations/sec (1 GPU, "replicas only")
roofline : dominant kernel = conv_tc (implicit-GEMM styled conv); achieved = algorithmic conv
           FLOPs / summed CUDA-event kernel time, against the MEASURED bf16 tensor peak
cpu_baseline / --impl reference: the CPU oracle port of the reference's PyTorch path
           (oracle/sg2_oracle.py; the Python reference itself cannot travel to the GPU box)
           timed on the host cores on a bounded sample (batch 2).

Multi-GPU: one process per GPU under torchrun; z batches are independent (weak scaling, no
data-path collective for image generation; one all-reduce for the covariance).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH = 32
COV_BATCH = 32
SIZE = 256
GFLOP_PER_IMG = 90.24          # algorithmic conv FLOPs of one 256^2 forward
METRIC = 'StyleGAN2-256 images/sec'


def conv_gflop_layers():
    """Algorithmic GFLOP per image of each styled conv (up layers counted on input res)."""
    chans = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128}
    out = {}
    out['layer2'] = 2 * 512 * 512 * 9 * 16 / 1e9
    n = 3
    cin = 512
    for res in (8, 16, 32, 64, 128, 256):
        cout = chans[res]
        out['layer%d' % n] = 2 * cin * cout * 9 * (res // 2) ** 2 / 1e9
        out['layer%d' % (n + 1)] = 2 * cout * cout * 9 * res ** 2 / 1e9
        n += 2
        cin = cout
    return out


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(tflops=float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1590.0))),
                    hbm=float(d.get('hbm_gbs', 6650.0)), source='measured (MEASURED_PEAKS.json, '
                    'sustained bf16 GEMM)')
    return dict(tflops=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(object):
    """Polls nvidia-smi (one-shot queries from a thread: its -lms loop block-buffers when
    piped) for SM clocks and throttle reasons while the timed region runs."""
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None

    def _poll(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                     '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
                for ln in out.stdout.strip().splitlines():
                    self.rows.append([p.strip() for p in ln.split(',')])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._thread = threading.Thread(target=self._poll, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=6)
        sm, mx, reasons = [], None, set()
        for parts in self.rows:
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx,
                    reasons=sorted(reasons), samples=len(sm))


def cpu_threads():
    """torch CPU convs at batch 2 stop scaling past ~16 threads (measured on the 128-core GPU
    host: 8 thr 1.11 s, 16 thr 0.96 s, 32 thr 0.98 s, 64 thr 1.41 s, 128 thr 19 s per forward),
    so the baseline uses the fastest setting, not the core count."""
    return max(1, min(os.cpu_count() or 1, 16))


def build_model(device):
    from oracle import sg2_oracle as orc      # only for the seeded-weights recipe (no compute)
    from rewriting_b200.utils.stylegan2 import SeqStyleGAN2
    model = orc.seeded_state_dict(lambda: SeqStyleGAN2(SIZE, style_dim=512, n_mlp=8, mconv='seq'))
    return model.to(device).eval()


def cpu_baseline_generator(seconds=12.0, batch=2):
    """Oracle port of the reference's generator forward on the host cores (bounded sample)."""
    from oracle import sg2_oracle as orc
    from rewriting_b200.utils.stylegan2 import SeqStyleGAN2
    from rewriting_b200.utils import zdataset
    cores = cpu_threads()
    torch.set_num_threads(cores)
    model = orc.seeded_state_dict(lambda: SeqStyleGAN2(SIZE, style_dim=512, n_mlp=8, mconv='seq'))
    sd = {k: v for k, v in model.state_dict().items()}
    z = zdataset.standard_z_sample(batch, 512, seed=1)
    with torch.no_grad():
        orc.generator_forward(sd, z)          # warm-up
        n, t0 = 0, time.time()
        while True:
            orc.generator_forward(sd, z)
            n += 1
            if time.time() - t0 > seconds or n >= 64:
                break
        dt = time.time() - t0
    out = dict(value=batch * n / dt, unit='images/s', cores=cores, kind='port',
               sample='%d forwards of batch %d (%.1f s), oracle/sg2_oracle.py generator_forward, '
                      'torch CPU fp32, %d threads' % (n, batch, dt, cores))
    # the other two quantities bench `extra` reports, on bounded samples:
    # key covariance (context forward to layer 8 + second moment) and the rewrite loop
    try:
        with torch.no_grad():
            t0 = time.time()
            zc = zdataset.standard_z_sample(4, 512, seed=1)
            keys = [orc.generator_forward(sd, zc[i:i + 2], upto_key_layer=8) for i in (0, 2)]
            orc.second_moment(keys)
            dt_cov = time.time() - t0
        k = keys[0][:1, :, 10:18, 12:21].contiguous()
        style = torch.ones(1, 512)
        w = sd['layer8.sconv.mconv.dconv.weight'].clone()
        tgt = orc.target_forward(k, style, w, sd['layer8.sconv.noise.weight'],
                                 sd['layer8.sconv.activate.bias']) * 1.5 + 0.3
        q, _ = torch.linalg.qr(torch.randn(512, 1))
        its = 10
        t0 = time.time()
        orc.insert_loop(w, k, style, tgt, sd['layer8.sconv.noise.weight'],
                        sd['layer8.sconv.activate.bias'], q.t().contiguous(), its)
        dt_ins = time.time() - t0
        out['extra'] = {'key_covariance_samples_per_s': 4 / dt_cov,
                        'insert_its_per_s': its / dt_ins,
                        'sample': '4 z to layer 8 + second moment (%.1f s); %d insert iterations on '
                                  'a 1x512x8x9 key (%.1f s); same oracle port, %d threads'
                                  % (dt_cov, its, dt_ins, cores)}
    except Exception as e:  # noqa: BLE001
        out['extra'] = {'error': '%s: %s' % (type(e).__name__, e)}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU path (oracle port) on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from oracle import sg2_oracle as orc
    from rewriting_b200.utils.stylegan2 import SeqStyleGAN2
    from rewriting_b200.utils import zdataset
    cores = cpu_threads()
    torch.set_num_threads(cores)
    model = orc.seeded_state_dict(lambda: SeqStyleGAN2(SIZE, style_dim=512, n_mlp=8, mconv='seq'))
    sd = dict(model.state_dict())
    sample = 2                              # images per step (bounded sample of the batch of 32)
    z = zdataset.standard_z_sample(sample, 512, seed=1)
    steps = max(1, min(args.steps, 40))
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 3))):
            orc.generator_forward(sd, z)
        t0 = time.time()
        for _ in range(steps):
            orc.generator_forward(sd, z)
        dt = time.time() - t0
    val = sample * steps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'SeqStyleGAN2-256 generator forward (mconv=seq), seeded random '
                               'weights, random z; each step = bounded sample of %d images of '
                               'the batch-32 workload' % sample},
        'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d steps x %d images, oracle port of the reference PyTorch '
                                   'path (the Python reference cannot travel to the GPU box)'
                                   % (steps, sample)},
        'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-extra', action='store_true', help='skip covariance / insert extras')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true',
                    help='time eager module calls instead of the CUDA-graph replay')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the product has no CPU path)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)

    from rewriting_b200 import _cabi, ops
    from rewriting_b200.utils import zdataset, nethook, runningstats
    from rewriting_b200 import dist as rdist
    _cabi.load()
    model = build_model(device)

    # per-rank z shard (weak scaling: every rank gets its own K+W batches of 32)
    n_batches = W + K
    z_all = zdataset.standard_z_sample(BATCH * n_batches * world, 512, seed=1)
    z_mine = z_all[rank * BATCH * n_batches:(rank + 1) * BATCH * n_batches].contiguous()
    z_dev = z_mine.to(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- per-kernel event timing of the dominant kernel (conv_tc) -------------------------
    # every conv_tc launch goes through one of these C-ABI entry points; bracket them with
    # CUDA events on the launching stream (argument positions: B, Cin, Cout, H, W)
    conv_events = []
    timing_on = {'on': False}
    CONV_ENTRY = {'rw_modconv_fwd': (10, 11, 12, 13, 14), 'rw_modconv_fwd_fused': (10, 11, 12, 13, 14),
                  'rw_modconv_up_fwd': (5, 6, 7, 8, 9), 'rw_modconv_up_fwd_cl': (5, 6, 7, 8, 9)}
    orig_call = _cabi.call

    def timed_call(name, *a):
        if not timing_on['on'] or name not in CONV_ENTRY:
            return orig_call(name, *a)
        iB, iCi, iCo, iH, iW = CONV_ENTRY[name]
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        orig_call(name, *a)
        ev1.record()
        conv_events.append((ev0, ev1, 2.0 * a[iB] * a[iCi] * a[iCo] * 9 * a[iH] * a[iW], 1))
    _cabi.call = timed_call

    # ---- public API objects: eager module and its CUDA-graph replay -------------------------
    from rewriting_b200.graphs import GraphedModule
    use_graph = not args.no_graph
    with torch.no_grad():
        model(z_dev[:BATCH])                               # one-off weight-plane preparation
        launches0 = _cabi.launch_count
        model(z_dev[:BATCH])
    launches_per_step = _cabi.launch_count - launches0     # kernels of ONE forward (mine only)
    runner = GraphedModule(model, z_dev[:BATCH]) if use_graph else model

    # ---- device-resident timing ------------------------------------------------------------
    with torch.no_grad():
        for i in range(W):
            runner(z_dev[i * BATCH:(i + 1) * BATCH])
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(W, W + K):
            flush.zero_()                                        # evict L2 between steps
            img = runner(z_dev[i * BATCH:(i + 1) * BATCH])
        e1.record()
        barrier()
        ms_dev = max_over_ranks(e0.elapsed_time(e1))
        launches = launches_per_step * K          # a graph replay launches the same kernels
        # per-kernel CUDA-event timing of the dominant kernel: eager replay of the same steps
        # (events cannot be read back from inside a graph), CPU running ahead of the GPU
        for i in range(2):
            model(z_dev[i * BATCH:(i + 1) * BATCH])
        torch.cuda.synchronize()
        timing_on['on'] = True
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(W, W + K):
            flush.zero_()
            model(z_dev[i * BATCH:(i + 1) * BATCH])
        t1.record()
        torch.cuda.synchronize()
        timing_on['on'] = False
        ms_eager = t0.elapsed_time(t1)
    conv_ms = sum(a.elapsed_time(b) for a, b, _, _ in conv_events)
    conv_flops = sum(f for _, _, f, _ in conv_events)
    conv_launches = sum(n for _, _, _, n in conv_events)
    conv_events.clear()

    # ---- end-to-end timing through the public API with host buffers ------------------------
    z_host = z_mine.pin_memory()
    out_host = torch.empty(BATCH, 3, SIZE, SIZE).pin_memory()
    out_hosts = [out_host, torch.empty(BATCH, 3, SIZE, SIZE).pin_memory()]
    with torch.no_grad():
        runner(z_host[:BATCH].to(device, non_blocking=True))
        barrier()
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(W, W + K):
            flush.zero_()
            if use_graph:       # pinned z -> static input (H2D), replay, images -> pinned host
                runner(z_host[i * BATCH:(i + 1) * BATCH], out=out_hosts[i & 1])
            else:
                zb = z_host[i * BATCH:(i + 1) * BATCH].to(device, non_blocking=True)
                out_host.copy_(model(zb), non_blocking=True)
        if use_graph:
            # the timed region ends when the LAST image batch has landed in host memory
            torch.cuda.current_stream().wait_stream(runner._copy_stream)
        s1.record()
        barrier()
        ms_e2e = max_over_ranks(s0.elapsed_time(s1))
    # SM clocks / throttle reasons sampled over all three timed loops (device-resident, per-kernel
    # replay, end-to-end): the first alone lasts ~0.2 s, i.e. one or two nvidia-smi polls
    clocks = sampler.stop() if rank == 0 else None

    value = BATCH * K * world / (ms_dev / 1e3)
    e2e_value = BATCH * K * world / (ms_e2e / 1e3)

    # ---- extras: covariance samples/s and rewrite-loop its/s --------------------------------
    extra = {}
    if not args.no_extra:
        from rewriting_b200 import fastpath
        with torch.no_grad():
            # what SeqStyleGanRewriter.collect_2nd_moment does per batch: generator up to layer
            # 8's conv (CUDA-graph replay), whose operand planes are the keys, then the col-GEMM
            key_runner = GraphedModule(lambda zz: fastpath.forward(model, zz, upto_key_layer=8),
                                       z_dev[:COV_BATCH])

            def cov_step(zb, r2m):
                planes = key_runner(zb)
                r2m.add_planes(planes.hi, planes.lo, planes.B * planes.H * planes.W)
            r2m = runningstats.RunningSecondMoment()
            for i in range(W):
                cov_step(z_dev[i * COV_BATCH:(i + 1) * COV_BATCH], r2m)
            r2m = runningstats.RunningSecondMoment()
            barrier()
            c0 = torch.cuda.Event(enable_timing=True)
            c1 = torch.cuda.Event(enable_timing=True)
            c0.record()
            for i in range(W, W + K):
                flush.zero_()
                cov_step(z_dev[i * COV_BATCH:(i + 1) * COV_BATCH], r2m)
            total = rdist.allreduce_moment_(r2m.mom2, r2m.count)     # the one collective
            c1.record()
            barrier()
            ms_cov = max_over_ranks(c0.elapsed_time(c1))
        extra['key_covariance_samples_per_s'] = COV_BATCH * K * world / (ms_cov / 1e3)
        extra['key_covariance'] = {'layer': 8, 'batch': COV_BATCH, 'rows_accumulated': total,
                                   'ms_per_step': ms_cov / K,
                                   'collective': 'one all_reduce(sum) of mom2[512,512] fp32 + '
                                                 'count after the last batch'}
        # every rank builds a rewriter: its constructor collects C collectively (sharded z +
        # one all-reduce) when torch.distributed is initialised; the edit itself is "replicas only"
        try:
            extra['insert_loop'] = bench_insert(model, z_dev, device)
        except Exception as e:  # noqa: BLE001
            extra['insert_loop'] = {'error': '%s: %s' % (type(e).__name__, e)}
        if world == 1:
            # BASELINE.json configs[1]: fused StyledConv forward + backward (dX, dstyle, dW, dbias,
            # dnoise), every layer shape of the 256^2 generator at batch 32
            try:
                del key_runner
                torch.cuda.empty_cache()
                from tools import bench_modconv
                r = bench_modconv.main(B=BATCH, quiet=True, save=False)
                sm = r['summary']
                extra['modconv_fwdbwd_b32'] = {
                    'fwd_ms': sm['total_fwd_ms'], 'fwdbwd_ms': sm['total_fwdbwd_ms'],
                    'fwd_TFLOPs': sm['fwd_TFLOPs'], 'fwdbwd_TFLOPs': sm['fwdbwd_TFLOPs'],
                    'per_layer_fwdbwd_ms': {l['layer']: round(l['fwdbwd_ms'], 4) for l in r['layers']},
                    'note': sm['note']}
            except Exception as e:  # noqa: BLE001
                extra['modconv_fwdbwd_b32'] = {'error': '%s: %s' % (type(e).__name__, e)}

    if rank == 0:
        peaks = measured_peaks()
        traffic = None
        prof = os.path.join(ROOT, 'profiles', 'r1_conv_tc_summary.json')
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    traffic = json.load(f).get('dram_bytes_per_launch')
            except Exception:
                traffic = None
        achieved = (conv_flops / 1e12) / (conv_ms / 1e3) if conv_ms > 0 else 0.0
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': ms_dev / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None,
            'dtype': 'bf16x3 (3-term split bf16 operands, fp32 accumulate; fp32 in/out)',
            'data': 'synthetic',
            'config': {'workload': 'SeqStyleGAN2-256 (mconv=seq, channel_multiplier=2) generator '
                                   'forward, batch=%d per GPU, seeded random weights, random z '
                                   '(zdataset seed 1)' % BATCH,
                       'execution': 'CUDA graph replay of model(z) (rewriting_b200.graphs.'
                                    'GraphedModule)' if use_graph else 'eager model(z)',
                       'global_batch': BATCH * world, 'parallelism': 'dp%d (independent z shards)'
                       % world, 'l2': 'flushed between steps (256 MiB memset, inside the timed region)',
                       'gflop_per_image': GFLOP_PER_IMG},
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tflops'],
                         'unit': 'TFLOP/s', 'frac': achieved / peaks['tflops'],
                         'traffic': traffic, 'kernel': 'rw::conv_tc_kernel<128> (all styled-conv '
                         'launches of the timed steps)', 'kernel_launches': conv_launches,
                         'kernel_ms_per_step': conv_ms / K,
                         'kernel_share_of_step': (conv_ms / K) / (ms_eager / K),
                         'timed_in': 'eager replay of the timed steps (%.2f ms/step); the headline '
                                     'value is the %s' % (ms_eager / K, 'CUDA-graph replay of the '
                                     'same module call' if use_graph else 'eager call'),
                         'peak_source': peaks['source'],
                         'note': 'algorithmic FLOPs (1x) against the cuBLAS-measured sustained bf16 '
                                 'peak; the 3-term split issues 3x the MMAs, so frac ~ 1/3 means the '
                                 'tensor pipe is as busy as in a cuBLAS GEMM; tensor-pipe utilisation '
                                 'per layer is in profiles/'},
            'e2e': {'value': e2e_value, 'unit': 'images/s', 'ms_per_step': ms_e2e / K,
                    'h2d_bytes_per_step': BATCH * 512 * 4,
                    'd2h_bytes_per_step': BATCH * 3 * SIZE * SIZE * 4},
            'gpu_launches': launches,
            'clocks': clocks,
            'extra': extra,
        }
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline_generator()
        elif world > 1:
            line['cpu_baseline'] = {'value': None, 'unit': 'images/s', 'cores': os.cpu_count(),
                                    'kind': 'port', 'sample': 'measured at N=1 only'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def bench_insert(model, z_dev, device, niter=400):
    """Rewrite-loop iterations/sec on a tight-paste sized key (config 4 shape: 1x512x8x9)."""
    import copy
    from rewriting_b200.rewrite import ganrewrite
    zds = torch.utils.data.TensorDataset(z_dev[:16].cpu())
    gw = ganrewrite.SeqStyleGanRewriter(copy.deepcopy(model), zds, 8)
    with torch.no_grad():
        bag = gw.context_model(gw.get_z(0))
        tgt = gw.target_model(bag)
        gin = type(bag)(bag, fmap=bag.fmap[:, :, 10:18, 12:21].contiguous())
        gout = type(bag)(bag, fmap=(tgt.fmap[:, :, 10:18, 12:21] * 1.5 + 0.3).contiguous())
        q, _ = torch.linalg.qr(torch.randn(512, 1, device=device))
        d = q.t().contiguous()
    gw.insert(gin, gout, d, niter=20)                         # warm-up
    ms = gw.insert(gin, gout, d, niter=niter, return_timing=True)
    bytes_per_iter = 6 * 512 * 512 * 9 * 4
    return {'its_per_s': niter / (ms / 1e3), 'ms_total': ms, 'niter': niter,
            'key_crop': [1, 512, 8, 9], 'rank': 1,
            'algorithmic_GBps': bytes_per_iter * niter / (ms / 1e3) / 1e9,
            'note': 'one rw_insert_loop launch for all iterations; W,m,v stay in smem/L2'}


if __name__ == '__main__':
    sys.exit(main())
