"""Host-side logic of the training path that needs no GPU: which chains take the native backward, the optimizer split and
its version-bump hook, cache invalidation on train() / eval(), and the loud failures on CPU tensors."""
import pytest
import torch

import licos_b200 as L
from licos_b200 import graphs


def test_native_backward_shape_rules():
    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
    assert net.g_a._native_backward_ok(torch.zeros(2, 3, 256, 256))
    assert net.g_a._native_backward_ok(torch.zeros(1, 3, 64, 96))
    assert not net.g_a._native_backward_ok(torch.zeros(1, 3, 250, 256))      # 250 / 2 = 125 is odd at the second layer
    assert not net.g_a._native_backward_ok(torch.zeros(1, 3, 255, 256))      # odd size at a stride-2 conv
    assert not net.g_a._native_backward_ok(torch.zeros(3, 256, 256))         # not a batch
    assert net.g_s._native_backward_ok(torch.zeros(2, 192, 7, 5))            # transposed convs double any size
    hyper = L.image_models["bmshj2018-hyperprior"](quality=1, pretrained=False)
    assert hyper.h_a._native_backward_ok(torch.zeros(2, 192, 16, 16)) and hyper.h_s._native_backward_ok(torch.zeros(2, 128, 4, 4))
    for bands in (1, 13):
        surg = L.get_model("bmshj2018-factorized", False, bands, 1)
        assert surg.g_a._native_backward_ok(torch.zeros(2, bands, 64, 64))
        assert surg.g_s._native_backward_ok(torch.zeros(2, 192, 4, 4)) == (bands <= 4)  # 13 bands out: no narrow last layer


def test_steps_follow_installed_modules():
    net = L.image_models["bmshj2018-factorized-relu"](quality=1, pretrained=False)
    kinds = [(type(m).__name__, epi) for m, _, epi, _ in net.g_a._steps()]
    assert [k for k, _ in kinds] == ["Conv2d"] * 4
    assert [e for _, e in kinds] == [3, 3, 3, 0]  # ReLU fused behind the first three convs, none behind the last


def test_optimizer_split_and_version_hook():
    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
    opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
    aux_params = [p for g in opt["aux"].param_groups for p in g["params"]]
    assert len(aux_params) == 1 and aux_params[0] is net.entropy_bottleneck.quantiles
    assert "fused" not in opt["net"].defaults or not opt["net"].defaults["fused"]  # CPU parameters: plain Adam
    w = net.g_a[0].weight
    w.grad = torch.zeros_like(w)
    v0 = w._version
    opt["net"].step()
    assert w._version > v0 + 0  # the in-place update and the post-step hook both bump it
    q = net.entropy_bottleneck.quantiles
    v1 = q._version
    opt["aux"].step()           # no gradient: Adam skips the parameter, the hook still invalidates caches keyed on it
    assert q._version > v1


def test_train_eval_switch_clears_kernel_layout_caches():
    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
    net.g_a._packed_cache[net.g_a[0]] = {"stale": object()}
    net.entropy_bottleneck._packed_key = ("stale",)
    net.train()
    assert len(net.g_a._packed_cache) == 0 and net.entropy_bottleneck._packed_key is None
    net.g_s._packed_cache[net.g_s[0]] = {"stale": object()}
    net.eval()
    assert len(net.g_s._packed_cache) == 0


def test_training_path_refuses_cpu_tensors():
    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False).train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.g_a(torch.rand(1, 3, 64, 64))
    crit = L.RateDistortionLoss()
    opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
    with pytest.raises(RuntimeError, match="no CPU path"):
        graphs.GraphedTrainStep(net, crit, opt, torch.rand(1, 3, 64, 64))


def test_reparametrizer_host_constants_match_buffers():
    g = L.GDN(8)
    for rp in (g.beta_reparam, g.gamma_reparam):
        assert rp.bound_f == float(rp.lower_bound.bound) and rp.pedestal_f == float(rp.pedestal)
