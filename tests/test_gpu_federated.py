"""-m gpu: the fused NCCL merge (licos_nccl_weighted_allreduce: prep kernel, ONE ncclAllReduce with a PreMulSum operator
whose scalar lives on the device, normalise kernel) through its C ABI on a one-rank communicator -- every piece of the
path runs (own communicator bootstrap through torch.distributed, device-scalar operator, spare-element normalisation);
the multi-rank arithmetic is covered on CPU by tests/test_federated_gloo.py and on 2 / 8 GPUs by bench.py's cfg-5 leg
(tools/federated_nccl_check.py compares the ranks bit for bit)."""
import socket

import pytest
import torch
import torch.distributed as dist

import licos_b200 as L
from licos_b200 import _lib
from licos_b200.federated import FlatState, federated_average

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_nccl_weighted_allreduce_single_rank(cuda):
    assert _lib.lib.licos_nccl_version() >= 21100  # PreMulSum exists since NCCL 2.11
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1, device_id=cuda)
    try:
        torch.manual_seed(1)
        net = L.get_model("bmshj2018-factorized", False, 3, 1).to(cuda)
        state = FlatState(net)
        before = state.flat.clone()
        x = torch.rand(1, 3, 64, 64, device=cuda)
        with torch.no_grad():
            ref_out = net.eval()(x)["x_hat"].clone()
        loss = torch.tensor([2.5], device=cuda)           # a device scalar: nothing is read back by the host
        federated_average(state, loss)
        torch.cuda.synchronize()
        assert torch.allclose(state.flat, before, rtol=3e-7, atol=0)   # (theta * u) / u
        assert float(state.buf[state.numel]) == pytest.approx(0.4)     # the spare element carried sum_r u_r = 1 / 2.5
        federated_average(state, 1.0, weights=[0.125])                 # explicit weights: exact powers of two
        federated_average(state, float("nan"))                         # vanishing, not poisonous
        torch.cuda.synchronize()
        assert torch.allclose(state.flat, before, rtol=1e-6, atol=0) and torch.isfinite(state.flat).all()
        with torch.no_grad():
            assert torch.allclose(net(x)["x_hat"], ref_out, atol=1e-2)  # the module computes on the merged views
        with pytest.raises(RuntimeError):
            net.half()
            federated_average(state, 1.0)
    finally:
        dist.destroy_process_group()
