"""NVTX ranges around the phases of the rewrite path (context / target / rendering forwards, key
second moment, key finding, insert loop), so that an nsys / ncu --nvtx timeline of a notebook
session reads in the paper's vocabulary (the reference has no tracing at all).
No-ops if the NVTX bindings are unavailable."""
import contextlib

try:
    import torch.cuda.nvtx as _nvtx
    _nvtx.range_push('rw:probe')
    _nvtx.range_pop()
except Exception:           # pragma: no cover
    _nvtx = None


@contextlib.contextmanager
def range(name):
    if _nvtx is None:
        yield
        return
    _nvtx.range_push(name)
    try:
        yield
    finally:
        _nvtx.range_pop()
