"""The body of /root/reference/licos/train.py:148-212 (``train_one_batch``: forward, rate-distortion loss, backward,
gradient clipping, main optimizer step, auxiliary loss + step) captured once as a CUDA graph and replayed per batch.

The eager loop costs ~360 kernel launches per step through Python; on a B200 the kernels of a 32-tile step take ~4 ms, the
launching ~5.5 ms.  A replay is one launch.  The step is the same code path as the eager loop (the modules' own
``forward`` / autograd nodes run during capture); nothing in it synchronises with the host or reads a device scalar.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn


class GraphedTrainStep:
    """``step = GraphedTrainStep(net, criterion, optimizers, example_batch, clip_max_norm)``; then per batch
    ``out = step(x)`` returns the dict of loss terms (device scalars that are overwritten by the next call).

    ``optimizers`` is the dict ``net_aux_optimizer`` returns (``{"net": ..., "aux": ...}``); on CUDA it builds fused,
    capturable Adam, which is what graph capture needs.  Batches must have the shape of ``example_batch``."""

    def __init__(self, net: nn.Module, criterion: nn.Module, optimizers: Dict[str, torch.optim.Optimizer],
                 example_batch: torch.Tensor, clip_max_norm: Optional[float] = 1.0, warmup: int = 3):
        if not example_batch.is_cuda:
            raise RuntimeError("licos_b200: GraphedTrainStep needs CUDA tensors on a B200 (no CPU path exists)")
        self.net, self.criterion, self.opt = net, criterion, optimizers
        self.clip = clip_max_norm
        self.x = example_batch.detach().clone()
        self._params = [p for p in net.parameters()]
        net.train()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # allocator, lazily created optimizer state, per-device kernel attributes
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        for o in self.opt.values():
            o.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.out = self._body()

    def _body(self):
        net, opt = self.net, self.opt
        opt["net"].zero_grad(set_to_none=True)
        opt["aux"].zero_grad(set_to_none=True)
        out = net(self.x)
        terms = self.criterion(out, self.x)
        terms["loss"].backward()
        if self.clip is not None and self.clip > 0:
            torch.nn.utils.clip_grad_norm_(net.parameters(), self.clip)
        opt["net"].step()
        aux = net.aux_loss()
        aux.backward()
        opt["aux"].step()
        res = {k: v.detach() for k, v in terms.items()}
        res["aux_loss"] = aux.detach()
        return res

    def __call__(self, x: torch.Tensor):
        if x.shape != self.x.shape:
            raise ValueError(f"batch shape {tuple(x.shape)} differs from the captured {tuple(self.x.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        # the replayed optimizer steps changed the parameters in place without touching their version counters:
        # bump them so the inference-side kernel-layout caches (keyed on versions) are rebuilt on next use
        for p in self._params:
            torch.autograd.graph.increment_version(p)
        return self.out


def _cached_tensors(net: nn.Module):
    """Every tensor the modules' kernel-layout caches hold right now (packed weights, GDN tables, bottleneck tables)."""
    keep = []

    def walk(v):
        if isinstance(v, torch.Tensor):
            keep.append(v)
        elif isinstance(v, (tuple, list)):
            for t in v:
                walk(t)
        elif isinstance(v, dict):
            for t in v.values():
                walk(t)
        elif hasattr(v, "tensors"):  # layers._Packed
            walk(v.tensors)

    for m in net.modules():
        cache = getattr(m, "_packed_cache", None)
        if cache is not None:
            for ent in list(cache.values()):
                walk(ent)
        ebp = getattr(m, "_packed", None)
        if ebp is not None:
            walk([getattr(ebp, n, None) for n in ("packed", "medians", "lut")])
    return keep


class GraphedForward:
    """``fwd = GraphedForward(net, example)``; then ``out = fwd(x)`` replays ``net(x)`` (eval mode, no autograd) as ONE CUDA
    graph launch.  For small inputs -- one crop per call, e.g. eval_utils.process_img -- the ~30 launches of a forward pass
    cost more on the host than on the GPU.  The parameters are constants of the graph (the packed weights and tables are
    the cached ones, kept alive here): after changing them (``load_state_dict``, a training step) build a new one.  Inputs
    must have the shape and dtype of ``example``; the returned tensors are overwritten by the next call."""

    def __init__(self, net: nn.Module, example: torch.Tensor, warmup: int = 3, fn=None):
        if not example.is_cuda:
            raise RuntimeError("licos_b200: GraphedForward needs CUDA tensors on a B200 (no CPU path exists)")
        from . import ops

        self.net = net.eval()
        self.x = example.detach().clone()
        self._fn = fn if fn is not None else (lambda t: net(t))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):  # fills the layout caches, the allocator, per-device kernel attributes
                self._fn(self.x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        ops._FROZEN_CAPTURE = True
        try:
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.out = self._fn(self.x)
        finally:
            ops._FROZEN_CAPTURE = False
        self._keep = _cached_tensors(net)  # the graph holds raw addresses of these: they must outlive it

    def __call__(self, x: torch.Tensor):
        if x.shape != self.x.shape or x.dtype != self.x.dtype:
            raise ValueError(f"input {tuple(x.shape)} {x.dtype} differs from the captured {tuple(self.x.shape)} {self.x.dtype}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out
