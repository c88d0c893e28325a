"""Federated weight merge (replaces /root/reference/licos/federation_utils.py:27-85).

The reference merges through a checkpoint file guarded by a lock file: one rank at a time does
``theta <- w * theta_local + (1 - w) * theta_central`` with ``w = best / (best + loss)`` over every state_dict
key (federation_utils.py:47-53) and adopts the result.  Here every rank is one GPU of one NVSwitch box, so the
merge is synchronous: all floating-point state lives in ONE flat fp32 buffer per rank (plus one spare element) and ONE
``ncclAllReduce`` with a PreMulSum operator moves it over NVLink (``licos_nccl_weighted_allreduce``): every rank's
buffer is multiplied by its un-normalised weight u_r = 1 / loss_r INSIDE the collective (device scalar, no host sync),
the spare element carries sum_r u_r, and one small kernel divides by it.  With two ranks and weights (w, 1 - w) this is
exactly the reference formula, which is what the parity tests check (tests/test_reference_glue.py runs the reference's
own federation_utils.py; tests/test_federated_gloo.py the 2-rank protocol on CPU).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, ops


def merge_pair(local_sd: Dict[str, torch.Tensor], central_sd: Dict[str, torch.Tensor], loss: float,
               best_loss: float) -> Dict[str, torch.Tensor]:
    """The reference's two-way rule, key by key, on the device the tensors live on."""
    w_local = best_loss / (best_loss + loss)
    w_central = loss / (best_loss + loss)
    out = {}
    for key, a in local_sd.items():
        b = central_sd[key].to(a.device)
        if a.is_cuda and a.dtype == torch.float32 and a.numel() > 0:
            out[key] = ops.weighted_sum2(a.contiguous(), b.contiguous(), w_local, w_central)
        else:  # integer tables (empty at train time in the reference) follow torch type promotion, as in the reference
            v = w_local * a
            v += w_central * b
            out[key] = v
    return out


class FlatState:
    """All floating-point parameters and buffers of a model re-homed into one contiguous fp32 buffer (the
    module's tensors become views), so a merge is one collective instead of ~150.

    Build it BEFORE anything captures addresses of the parameters (``GraphedTrainStep``, optimizers with fused state are
    fine: they hold the Parameter objects, whose ``.data`` is re-pointed here, but a CUDA graph replays raw addresses) and do
    not ``.to()`` / ``.float()`` the module afterwards: both would leave the graph or the module training storage the merge
    never touches.  ``verify()`` (called by every merge) checks that the views are still in place."""

    def __init__(self, net: nn.Module):
        self.net = net
        tensors: List[torch.Tensor] = []
        seen = set()
        for t in list(net.parameters()) + list(net.buffers()):
            if t.dtype == torch.float32 and t.numel() > 0 and id(t) not in seen:
                seen.add(id(t))
                tensors.append(t)
        total = sum(t.numel() for t in tensors)
        device = tensors[0].device
        # + the spare element that carries the sum of the weights through the collective (padded to a 16-byte multiple)
        self.buf = torch.zeros((total + 1 + 3) // 4 * 4, dtype=torch.float32, device=device)
        self.flat = self.buf[:total]
        off = 0
        with torch.no_grad():
            for t in tensors:
                n = t.numel()
                view = self.flat[off:off + n].view(t.shape)
                view.copy_(t)
                t.data = view
                off += n
        self.numel = total
        self._tensors = tensors
        self._scalar = torch.zeros(1, dtype=torch.float32, device=device)

    def verify(self, full: bool = False) -> None:
        """Every merge checks the first, the last and one middle tensor (a move / re-type re-homes all of them at once);
        ``full=True`` walks them all."""
        lo = self.flat.data_ptr()
        hi = lo + 4 * self.numel
        ts = self._tensors
        for t in (ts if full or len(ts) < 4 else (ts[0], ts[len(ts) // 2], ts[-1])):
            if not (lo <= t.data_ptr() < hi):
                raise RuntimeError("licos_b200.FlatState: a parameter no longer lives in the flat buffer (the module was moved "
                                   "or re-typed after FlatState was built); rebuild FlatState -- and any CUDA graph -- first")

    def touch(self) -> None:
        """Bump parameter versions so kernel-layout weight caches are rebuilt after an in-place merge."""
        for p in self.net.parameters():
            torch.autograd.graph.increment_version(p)  # no kernel launch (p.add_(0) cost ~150 launches per merge)


class NcclMerger:
    """This package's own NCCL communicator over the ranks of a torch.distributed group (the 128-byte unique id travels
    through that group once), used by ``licos_nccl_weighted_allreduce``.  One per process."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import ctypes

        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            _lib.check(_lib.lib.licos_nccl_unique_id(ident.data_ptr()), "nccl_unique_id")
        box = [ident.numpy().tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._id = (ctypes.c_char * 128).from_buffer_copy(box[0])
        comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib.licos_nccl_comm_create(ctypes.addressof(self._id), self.world, self.rank, ctypes.byref(comm)),
                       "nccl_comm_create")
        self.comm, self.device = comm, device

    def merge(self, state: FlatState, loss: Optional[torch.Tensor] = None, weight: Optional[torch.Tensor] = None) -> None:
        """``loss`` / ``weight``: one-element fp32 DEVICE tensors (this rank's validation loss, or its explicit weight)."""
        t = weight if weight is not None else loss
        if t is None or not t.is_cuda or t.dtype != torch.float32:
            raise ValueError("NcclMerger.merge needs a float32 CUDA scalar")
        _lib.check(_lib.lib.licos_nccl_weighted_allreduce(
            self.comm, state.buf.data_ptr(), state.numel, None if weight is not None else t.data_ptr(),
            t.data_ptr() if weight is not None else None, state._scalar.data_ptr(), ops._stream()), "nccl_weighted_allreduce")

    def close(self) -> None:
        if self.comm:
            _lib.lib.licos_nccl_comm_destroy(self.comm)
            self.comm = None


_MERGERS = {}


def _device_scalar(v, device) -> torch.Tensor:
    if isinstance(v, torch.Tensor):
        return v.detach().reshape(1).to(device=device, dtype=torch.float32)
    return torch.tensor([float(v)], dtype=torch.float32).to(device, non_blocking=True)


def federated_average(state: FlatState, loss, group: Optional[dist.ProcessGroup] = None,
                      weights: Optional[Sequence[float]] = None) -> torch.Tensor:
    """Synchronous N-way merge.  ``loss``: this rank's loss (float or tensor; a device tensor avoids any host traffic).
    ``weights`` (one per rank) default to the reference's rule generalised to N ranks: w_i proportional to 1 / loss_i,
    normalised by their sum; a rank whose loss is not a positive finite number gets a vanishing weight."""
    state.verify()
    rank = dist.get_rank(group)
    flat, n = state.flat, state.numel
    if flat.is_cuda and dist.get_backend(group) == "nccl":
        key = id(group)
        if key not in _MERGERS:
            _MERGERS[key] = NcclMerger(flat.device, group)
        if weights is None:
            _MERGERS[key].merge(state, loss=_device_scalar(loss, flat.device))
        else:
            _MERGERS[key].merge(state, weight=_device_scalar(weights[rank], flat.device))
    else:
        # the same protocol in torch ops: gloo tests of the host logic on CPU ranks
        if weights is None:
            l = float(loss)
            u = 1.0 / max(l, 1e-12) if (l == l and l > 0 and l != float("inf")) else 1e-20
        else:
            u = float(weights[rank])
        with torch.no_grad():
            flat.mul_(u)
            state.buf[n] = u
            dist.all_reduce(state.buf[:n + 1], op=dist.ReduceOp.SUM, group=group)
            flat.div_(state.buf[n])
    state.touch()
    return flat
