"""CUDA-graph capture of a fixed-shape forward (launch-bound inner loops).

One generator forward at batch 32 is ~190 kernel launches (31 tensor-core convs plus the
prep / blur / ToRGB / mapping kernels); captured once, a replay costs one launch.  The
captured region is exactly the public module call — `GraphedModule(model, z)(z_new)` returns
what `model(z_new)` returns — and is valid while the module's parameters keep their storage
(in-place edits such as the rewriter's are picked up; the bf16 weight planes are refreshed
after an edit: `__call__` compares the `_version` of every parameter of the wrapped module
with the versions seen at capture and re-captures by itself when one changed — in-place edits
by the rewriter, `load_state_dict`, optimizer steps all bump it).
"""
import torch


class GraphedModule(object):
    def __init__(self, module, example_input, warmup=3, parameters=None):
        if not example_input.is_cuda:
            raise RuntimeError('GraphedModule needs a CUDA example input')
        self.module = module
        # callable returning the parameters whose `_version` invalidates the capture (defaults
        # to module.parameters() for an nn.Module; pass it for plain functions / lambdas)
        self._parameters = parameters
        self.static_in = example_input.detach().clone()
        self.graph = None
        self.static_out = None
        self._warmup = warmup
        self._copy_stream = torch.cuda.Stream()
        self._stage = None          # two device staging buffers for pipelined D2H
        self._stage_free = [None, None]
        self._turn = 0
        self._versions = None
        self._keepalive = None
        self.refresh()

    def _param_versions(self):
        if self._parameters is not None:
            params = self._parameters()
        elif isinstance(self.module, torch.nn.Module):
            params = self.module.parameters()
        else:
            params = ()
        return tuple((p.data_ptr(), p._version) for p in params)

    def refresh(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self._warmup):       # fills the noise / weight-plane caches
                self.module(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self.module(self.static_in)
        self._versions = self._param_versions()
        from . import ops
        self._keepalive = ops.cached_device_state()

    def __call__(self, x, out=None):
        """Replays the forward on `x` (any device; copied into the static input).  Returns the
        static output tensor (overwritten by the next call).  With a pinned host tensor `out`,
        the result is moved to the host on a side stream through two staging buffers, so the
        device->host copy of step i overlaps the compute of step i+1; call `sync()` before
        reading `out`."""
        if self._param_versions() != self._versions:
            # a parameter was edited (or re-loaded) since the capture: the bf16 weight planes the
            # graph reads are stale — capture again against the current weights
            self.refresh()
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        if out is None:
            return self.static_out
        if out.is_cuda:
            out.copy_(self.static_out, non_blocking=True)
            return out
        if self._stage is None:
            self._stage = [torch.empty_like(self.static_out) for _ in range(2)]
        i = self._turn
        self._turn ^= 1
        main = torch.cuda.current_stream()
        if self._stage_free[i] is not None:           # the copy that last used this buffer
            main.wait_event(self._stage_free[i])
        self._stage[i].copy_(self.static_out, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            out.copy_(self._stage[i], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self._stage_free[i] = done
        return out

    def last_copy_event(self):
        """CUDA event recorded after the device->host copy the latest `__call__(x, out=host)`
        issued (None if it returned device data)."""
        return self._stage_free[self._turn ^ 1]

    def sync(self):
        """Wait for outstanding device->host copies issued by __call__(x, out=host)."""
        self._copy_stream.synchronize()
