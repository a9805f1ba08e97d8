"""Bulk image generation and its output side (BASELINE config 5).

Reference loops being replaced:
  * utils/get_samples.py:114-129 — `nimgs // 10 + 1` batches of 10; batch j draws
    `z = z_sample_for_model(g, 10, seed=len(samples))` (= seed 10*j), runs the generator and moves
    every image to the CPU one by one into a Python list (50 010 images for nimgs = 50 000);
  * metrics/sample.py:19-37 (also sample_edited.py:55-60, make_watermark_images.py:99-131) —
    batch 1, `z = z_sample_for_model(model, 1, seed=imgnum + offset)`, `.cpu()`, PIL PNG through
    a SaveImagePool.

Same z (bit-identical numpy streams) and the same noise per image here — sample i of a reference
batch of 10 sees row i of `RandomState(0).randn(10, H*W)`, a batch-1 image always row 0, which a
`noise_period` table reproduces for any number of reference batches per pass — but one CUDA-graph
replay covers `group` reference batches, the last ToRGB combine writes NHWC uint8 directly
(`out_dtype=torch.uint8`: 4x less PCIe traffic, no fp32 image in HBM), the device->host copy of
pass n overlaps the compute of pass n+1, and with torch.distributed initialised rank r of R takes
the reference batches j = r (mod R) — independent units, no exchange.
"""
import os
import threading

import numpy
import torch

from . import dist as rdist
from . import fastpath
from .graphs import GraphedModule
from .utils import zdataset


def z_for_batch(j, batch=10, depth=512):
    """z of batch j exactly as the reference draws it (seed = number of samples so far)."""
    return zdataset.standard_z_sample(batch, depth, seed=batch * j)


def to_uint8_nhwc(images):
    """[B,3,H,W] in [-1,1] -> [B,H,W,3] uint8, clamp(x*127.5+127.5, 0, 255) truncated (torch ops;
    the fast path's `out_u8` kernel computes the same bytes without the fp32 image)."""
    return (images * 127.5 + 127.5).clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def _runner(model, z0, out_dtype, period):
    """CUDA-graph replay of one pass of z0.shape[0] images (noise rows repeat with `period`)."""
    if fastpath.eligible(model, z0):
        if out_dtype == torch.uint8:
            fwd = lambda z: fastpath.forward(model, z, noise_period=period, out_u8=True)
        else:
            fwd = lambda z: fastpath.forward(model, z, noise_period=period)
    else:                         # hooked / edited module trees: child by child, reference batches
        def fwd(z):
            outs = [model(z[i:i + period]) for i in range(0, z.shape[0], period)]
            img = torch.cat(outs)
            return to_uint8_nhwc(img) if out_dtype == torch.uint8 else img
    return GraphedModule(fwd, z0, parameters=model.parameters)


def _pump(runner, passes, host, zdim, sink=None):
    """Drives `runner` over `passes` (callables returning the pass's z on the CPU) with double
    buffered pinned z staging and pipelined D2H into `host[n % len(host)]`.  With a `sink`, the
    host buffers are a ring: `sink(n, host_view)` is called once pass n has landed, before its
    buffer is reused (len(host) >= 3 keeps two passes in flight)."""
    zpin = torch.empty((2,) + tuple(zdim), dtype=torch.float32).pin_memory()
    consumed = [None, None]       # event after the H2D copy that last read each pinned z slot
    landed = []                   # (pass index, copy-done event) not yet handed to the sink
    ring = len(host)

    def drain(keep):
        while len(landed) > keep:
            n, ev = landed.pop(0)
            ev.synchronize()
            sink(n, host[n % ring])
    with torch.no_grad():
        for n, make_z in enumerate(passes):
            slot = n & 1
            if consumed[slot] is not None:
                consumed[slot].synchronize()
            zpin[slot].copy_(make_z())
            if sink is not None:
                drain(ring - 1)               # the buffer about to be overwritten is free again
            runner(zpin[slot], out=host[n % ring])
            if sink is not None:
                landed.append((n, runner.last_copy_event()))
            consumed[slot] = torch.cuda.Event()
            consumed[slot].record()
        runner.sync()
        torch.cuda.synchronize()
        if sink is not None:
            drain(0)


def get_samples(model, nimgs=50000, batch=10, out_dtype=torch.float32, shard=True,
                reference_count=True, group=4, sink=None):
    """images on the CPU (pinned) — [n,3,H,W] fp32, or [n,H,W,3] uint8 for out_dtype=torch.uint8 —
    and the list of global reference-batch indices they came from.

    reference_count=True reproduces the reference's `nimgs // batch + 1` batches
    (50 010 images for nimgs = 50 000); False generates
    ceil(nimgs / batch) batches.  `group` reference batches run per graph replay.
    `sink(images, batch_indices)`: stream the passes to a consumer (e.g. `ImageWriter.add`)
    through a ring of three pinned buffers instead of keeping all images in host memory
    (50 010 uint8 images are 9.8 GB); the function then returns (None, batch indices)."""
    device = next(model.parameters()).device
    nb = nimgs // batch + 1 if reference_count else -(-nimgs // batch)
    R, r = (rdist.world_size(), rdist.rank()) if shard else (1, 0)
    mine = [j for j in range(nb) if j % R == r]
    if not mine:
        return torch.empty(0), []
    group = max(1, min(group, len(mine)))
    chunks = [mine[i:i + group] for i in range(0, len(mine), group)]
    full = [c for c in chunks if len(c) == group]
    tail = [c for c in chunks if len(c) != group]
    zdepth = getattr(model, 'style_dim', 512)
    z0 = torch.cat([z_for_batch(j, batch, zdepth) for j in chunks[0]]).to(device)
    runner = _runner(model, z0, out_dtype, batch)
    shape = tuple(runner.static_out.shape[1:])
    per = group * batch
    passes = [(lambda c=c: torch.cat([z_for_batch(j, batch, zdepth) for j in c])) for c in full]
    if sink is not None:
        ringbuf = [torch.empty((per,) + shape, dtype=out_dtype).pin_memory() for _ in range(3)]
        _pump(runner, passes, ringbuf, z0.shape, sink=lambda n, v: sink(v, full[n]))
        host = None
    else:
        host = torch.empty((len(mine) * batch,) + shape, dtype=out_dtype).pin_memory()
        _pump(runner, passes, [host[n * per:(n + 1) * per] for n in range(len(full))], z0.shape)
    for c in tail:                 # ragged last pass: its own (smaller) replay
        zt = torch.cat([z_for_batch(j, batch, zdepth) for j in c]).to(device)
        rt = _runner(model, zt, out_dtype, batch)
        dst = host[len(full) * per:] if host is not None else \
            torch.empty((len(c) * batch,) + shape, dtype=out_dtype).pin_memory()
        rt(zt, out=dst)
        rt.sync()
        torch.cuda.synchronize()
        if sink is not None:
            sink(dst, c)
    return host, mine


def sample_images(model, imgnums, offset=0, out_dtype=torch.uint8, shard=True, group=32):
    """The `metrics/sample.py` loop: image `imgnum` is the generator at
    `z_sample_for_model(model, 1, seed=imgnum + offset)` run as a batch of ONE (so every image
    sees noise row 0).  `group` of them run per replay with a period-1 noise table.  Returns
    (images on the CPU, the imgnums of this rank)."""
    device = next(model.parameters()).device
    imgnums = list(imgnums)
    R, r = (rdist.world_size(), rdist.rank()) if shard else (1, 0)
    lo, hi = rdist.shard_range(len(imgnums), r, R)
    mine = imgnums[lo:hi]
    if not mine:
        return torch.empty(0), []
    zdepth = getattr(model, 'style_dim', 512)

    def z_of(nums):
        return torch.cat([zdataset.standard_z_sample(1, zdepth, seed=n + offset) for n in nums])
    group = max(1, min(group, len(mine)))
    nfull = len(mine) // group
    z0 = z_of(mine[:group]).to(device)
    runner = _runner(model, z0, out_dtype, 1)
    shape = tuple(runner.static_out.shape[1:])
    host = torch.empty((len(mine),) + shape, dtype=out_dtype).pin_memory()
    views = [host[n * group:(n + 1) * group] for n in range(nfull)]
    _pump(runner, [(lambda n=n: z_of(mine[n * group:(n + 1) * group])) for n in range(nfull)],
          views, z0.shape)
    if nfull * group < len(mine):
        zt = z_of(mine[nfull * group:]).to(device)
        rt = _runner(model, zt, out_dtype, 1)
        rt(zt, out=host[nfull * group:])
        rt.sync()
        torch.cuda.synchronize()
    return host, mine


# ------------------------------------------------------------------------------------------
# writers (reference: utils/imgsave.py SaveImagePool + torchvision ToPILImage, metrics/sample.py:33-37)
# ------------------------------------------------------------------------------------------
class ImageWriter(object):
    """Writes NHWC uint8 images as `<dirname>/<imgnum>.png` on worker threads (the reference's
    SaveImagePool pattern) or as one `images.npz` (uint8 array `images`, int64 `imgnums`)."""

    def __init__(self, dirname, fmt='png', workers=8):
        assert fmt in ('png', 'npz')
        self.dirname, self.fmt = dirname, fmt
        os.makedirs(dirname, exist_ok=True)
        self._threads = []
        self._sem = threading.Semaphore(workers)
        self._npz = ([], [])

    def add(self, images_u8_nhwc, imgnums):
        arr = images_u8_nhwc.numpy() if isinstance(images_u8_nhwc, torch.Tensor) else images_u8_nhwc
        assert arr.dtype == numpy.uint8 and arr.ndim == 4 and arr.shape[3] == 3
        if self.fmt == 'npz':
            self._npz[0].append(numpy.array(arr))
            self._npz[1].extend(int(n) for n in imgnums)
            return
        arr = numpy.array(arr)          # detach from the pinned staging buffer

        def work(a, nums):
            from PIL import Image
            try:
                for img, n in zip(a, nums):
                    Image.fromarray(img, 'RGB').save(os.path.join(self.dirname, '%d.png' % n),
                                                     optimize=False, compress_level=1)
            finally:
                self._sem.release()
        self._sem.acquire()
        t = threading.Thread(target=work, args=(arr, list(imgnums)), daemon=True)
        t.start()
        self._threads.append(t)

    def join(self):
        for t in self._threads:
            t.join()
        self._threads = []
        if self.fmt == 'npz' and self._npz[0]:
            numpy.savez(os.path.join(self.dirname, 'images.npz'),
                        images=numpy.concatenate(self._npz[0]),
                        imgnums=numpy.array(self._npz[1], dtype=numpy.int64))
            self._npz = ([], [])


# ------------------------------------------------------------------------------------------
# Frechet statistics (reference: metrics/fid.py:137-187, numpy + scipy.linalg.sqrtm on the host
# after a TensorFlow Inception pass).  The statistics half is torch-native here; the Inception
# network itself needs downloaded weights and stays pluggable (`features` is any [N, D] tensor).
# ------------------------------------------------------------------------------------------
def pt_to_float255_nhwc(images):
    """metrics/fid.py:178-181 `pt_to_np`: [-1,1] NCHW -> [0,255] float NHWC, on the device."""
    return ((images / 2 + 0.5) * 255).clamp(0, 255).permute(0, 2, 3, 1).contiguous()


def activation_statistics(features):
    """(mu [D], sigma [D,D]) = (mean, np.cov(rowvar=False)) of [N, D] features, fp64 on the
    features' device."""
    f = features.double()
    mu = f.mean(0)
    c = f - mu
    return mu, c.t() @ c / (f.shape[0] - 1)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """||mu1-mu2||^2 + Tr(S1 + S2 - 2 sqrt(S1 S2))  (metrics/fid.py:137-175).  Tr sqrt(S1 S2) is
    taken from the symmetric form sqrt(S1) S2 sqrt(S1) (same eigenvalues as S1 S2, but a
    symmetric PSD matrix: two eigh calls instead of a general matrix square root)."""
    mu1, mu2 = mu1.double(), mu2.double()
    s1, s2 = sigma1.double(), sigma2.double()

    def tr_sqrt_prod(a, b):
        w, v = torch.linalg.eigh(a)
        ra = (v * w.clamp_min(0).sqrt()) @ v.t()
        m = ra @ b @ ra
        ev = torch.linalg.eigvalsh((m + m.t()) / 2)
        return ev.clamp_min(0).sqrt().sum()
    t = tr_sqrt_prod(s1, s2)
    if not torch.isfinite(t):
        off = torch.eye(s1.shape[0], dtype=s1.dtype, device=s1.device) * eps
        t = tr_sqrt_prod(s1 + off, s2 + off)
    diff = mu1 - mu2
    return float(diff.dot(diff) + s1.trace() + s2.trace() - 2 * t)
