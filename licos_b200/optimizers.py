"""compressai.optimizers.net_aux_optimizer (reached from /root/reference/licos/utils.py:65-73): every parameter
whose name ends in ``.quantiles`` goes to the auxiliary optimizer, everything else to the main one."""
from __future__ import annotations

from typing import Any, Dict

import torch
import torch.nn as nn


def _bump_versions(optimizer, args, kwargs) -> None:
    """Fused / foreach steps change parameters in place without bumping their version counters; the kernel-layout weight
    caches of the inference path are keyed on those counters."""
    for group in optimizer.param_groups:
        for p in group["params"]:
            torch.autograd.graph.increment_version(p)


def net_aux_optimizer(net: nn.Module, conf: Dict[str, Dict[str, Any]]) -> Dict[str, torch.optim.Optimizer]:
    named = dict(net.named_parameters())
    groups = {"net": [], "aux": []}
    for name in sorted(named):
        if named[name].requires_grad:
            groups["aux" if name.endswith(".quantiles") else "net"].append(named[name])
    out = {}
    for key in ("net", "aux"):
        kwargs = dict(conf[key])
        kind = kwargs.pop("type")
        if kind in ("Adam", "AdamW") and "fused" not in kwargs and "foreach" not in kwargs and groups[key] and all(
                p.is_cuda for p in groups[key]):
            kwargs["fused"] = True  # same update rule, one multi-tensor launch instead of one per parameter
            kwargs.setdefault("capturable", True)  # step counters on the device: no host sync, CUDA-graph safe
        out[key] = getattr(torch.optim, kind)(groups[key], **kwargs)
        out[key].register_step_post_hook(_bump_versions)
    return out
