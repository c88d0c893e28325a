"""Host logic of the CompressAI-mirroring modules: construction, surgery, state_dict contract, update() tables,
error behaviour.  CPU only -- anything that would compute on the hot path must raise without a GPU."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import licos_b200 as L
from licos_b200 import synth
from oracle import compressai_ref as R


@pytest.mark.parametrize("name,q", [("bmshj2018-factorized", 1), ("bmshj2018-factorized-relu", 1),
                                    ("bmshj2018-hyperprior", 1), ("bmshj2018-hyperprior", 6)])
def test_state_dict_contract_matches_oracle(name, q):
    torch.manual_seed(0)
    net = L.image_models[name](quality=q, pretrained=False)
    ref = R.image_models[name](quality=q)
    sd, rsd = net.state_dict(), ref.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    assert all(sd[k].shape == rsd[k].shape and sd[k].dtype == rsd[k].dtype for k in sd)
    ref.load_state_dict(sd)  # strict
    net.load_state_dict(rsd)
    # arithmetic on every key is legal (federation_utils.py:51-53)
    for k in sd:
        _ = 0.3 * sd[k] + 0.7 * rsd[k]


@pytest.mark.parametrize("in_ch,count", [(1, 2980737), (3, 2998147), (13, 3108237)])
def test_get_model_surgery(in_ch, count):
    net = L.get_model("bmshj2018-factorized", False, in_ch, 1)
    assert sum(p.numel() for p in net.parameters()) == count
    assert net.g_a[0].in_channels == in_ch and net.g_s[6].out_channels == in_ch
    assert net.entropy_bottleneck.filters == (in_ch, in_ch, 3, 3) and net.entropy_bottleneck.channels == 192
    assert isinstance(net.g_a[0], nn.Conv2d) and net.g_a[0].stride == (2, 2) and net.g_a[0].padding == (2, 2)
    assert net.g_s[6].output_padding == (1, 1)
    with pytest.raises(ValueError):
        L.get_model("mbt2018", False, 1, 1)
    with pytest.raises(ValueError):
        L.image_models["bmshj2018-factorized"](quality=9)
    with pytest.raises(RuntimeError):
        L.image_models["bmshj2018-factorized"](quality=1, pretrained=True)


@pytest.mark.parametrize("in_ch", [1, 3, 13])
def test_update_tables_bit_exact_vs_oracle(in_ch):
    torch.manual_seed(42)
    net = L.get_model("bmshj2018-factorized", False, in_ch, 1)
    synth.condition_weights(net)
    ref = R.get_model("bmshj2018-factorized", False, in_ch, 1)
    ref.load_state_dict(net.state_dict())
    assert net.update() is True and net.update() is False and net.update(force=True) is True
    ref.update(force=True)
    for name in ("_quantized_cdf", "_cdf_length", "_offset"):
        a, b = getattr(net.entropy_bottleneck, name), getattr(ref.entropy_bottleneck, name)
        assert a.dtype == torch.int32 and torch.equal(a, b), name
    # updated checkpoints reload (buffer resizing), also under a {"state_dict": ...} wrapper
    ckpt = {"state_dict": net.state_dict()}
    fresh = L.get_model("bmshj2018-factorized", False, in_ch, 1)
    fresh.load_state_dict(ckpt["state_dict"])
    assert torch.equal(fresh.entropy_bottleneck._quantized_cdf, net.entropy_bottleneck._quantized_cdf)


def test_hyperprior_tables_bit_exact_vs_oracle():
    torch.manual_seed(42)
    net = L.image_models["bmshj2018-hyperprior"](quality=1)
    synth.condition_weights(net)
    ref = R.image_models["bmshj2018-hyperprior"](quality=1)
    ref.load_state_dict(net.state_dict())
    net.update()
    ref.update()
    for mod in ("entropy_bottleneck", "gaussian_conditional"):
        for name in ("_quantized_cdf", "_cdf_length", "_offset"):
            assert torch.equal(getattr(getattr(net, mod), name), getattr(getattr(ref, mod), name)), (mod, name)
    assert torch.equal(net.gaussian_conditional.scale_table, ref.gaussian_conditional.scale_table)


def test_newer_parameterlist_names_are_accepted():
    net = L.image_models["bmshj2018-factorized"](quality=1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    renamed = {}
    for k, v in sd.items():
        k2 = k
        for old, new in (("_matrix", "matrices."), ("_bias", "biases."), ("_factor", "factors.")):
            if f"entropy_bottleneck.{old}" in k:
                k2 = k.replace(f"entropy_bottleneck.{old}", f"entropy_bottleneck.{new}")
        renamed[k2] = v + 1 if v.dtype == torch.float32 and v.numel() > 0 else v
    net.load_state_dict(renamed)
    assert torch.equal(net.entropy_bottleneck._matrix0, sd["entropy_bottleneck._matrix0"] + 1)


def test_aux_loss_optimizer_split_and_rd_loss_shape():
    torch.manual_seed(1)
    net = L.image_models["bmshj2018-hyperprior"](quality=1)
    ref = R.image_models["bmshj2018-hyperprior"](quality=1)
    ref.load_state_dict(net.state_dict())
    assert torch.allclose(net.aux_loss(), ref.aux_loss())
    opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
    aux = [p for g in opt["aux"].param_groups for p in g["params"]]
    assert len(aux) == 1 and aux[0] is net.entropy_bottleneck.quantiles
    n_net = sum(p.numel() for g in opt["net"].param_groups for p in g["params"])
    assert n_net + aux[0].numel() == sum(p.numel() for p in net.parameters())
    assert opt["net"].param_groups[0]["lr"] == 1e-4 and opt["aux"].param_groups[0]["lr"] == 1e-3


def test_no_cpu_path_and_error_behaviour():
    net = L.image_models["bmshj2018-hyperprior"](quality=1).eval()
    x = torch.rand(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU"):
        net(x)
    with pytest.raises(RuntimeError, match="no CPU"):
        net.g_a(x)
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            net.entropy_bottleneck(torch.rand(1, 128, 4, 4))
        with pytest.raises(RuntimeError):
            net.gaussian_conditional.build_indexes(torch.rand(1, 192, 4, 4))
    eb = L.EntropyBottleneck(8)
    with pytest.raises(ValueError, match="update"):
        eb.compress(torch.zeros(1, 8, 2, 2))
    with pytest.raises(ValueError):
        eb.quantize(torch.zeros(2), "bogus")
    with pytest.raises(ValueError):
        L.GaussianConditional([3.0, 1.0])
    with pytest.raises(ValueError):
        L.EntropyBottleneck(4, filters=(32, 3))
    with pytest.raises(NotImplementedError):
        from licos_b200.layers import _conv_kind
        _conv_kind(nn.Conv2d(3, 3, 7, 2, 3))


def test_gdn_reparam_and_lower_bound_gradient():
    g, rg = L.GDN(4), R.GDN(4)
    rg.load_state_dict(g.state_dict())
    x = torch.randn(2, 4, 3, 3)
    # the differentiable module expression equals the oracle's (used only under autograd on a GPU)
    assert torch.allclose(L.GDN.forward(g, x), rg(x))
    xb = torch.tensor([0.5, 2.0, 0.5], requires_grad=True)
    L.LowerBound(1.0)(xb).backward(torch.tensor([1.0, 1.0, -1.0]))
    assert xb.grad.tolist() == [0.0, 1.0, -1.0]


def test_update_follows_the_forward_likelihood_form():
    """One switch drives forward() and update(), as in every upstream release: the tables update() writes describe the
    density forward() evaluates.  The two forms round differently on some rows, so the switch must actually reach update()."""
    import torch
    from oracle import compressai_ref as R

    torch.manual_seed(7)
    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
    from licos_b200 import synth
    synth.condition_weights(net)
    ref = R.image_models["bmshj2018-factorized"](quality=1)
    ref.load_state_dict(net.state_dict())
    tables = {}
    for form in ("plain", "stable"):
        net.entropy_bottleneck.likelihood_form = ref.entropy_bottleneck.likelihood_form = form
        net.update(force=True)
        ref.update(force=True)
        assert torch.equal(net.entropy_bottleneck._quantized_cdf, ref.entropy_bottleneck._quantized_cdf), form
        tables[form] = net.entropy_bottleneck._quantized_cdf.clone()
    assert tables["plain"].shape == tables["stable"].shape
    assert not torch.equal(tables["plain"], tables["stable"])  # the forms are not interchangeable: hence one switch


def test_kernel_parameter_cache_is_not_part_of_a_copy_or_pickle():
    """EntropyBottleneck caches a ctypes parameter block after the first kernel call; deepcopy (EMA / best-model
    snapshots) and pickling must drop it, not choke on it."""
    import copy
    import io
    import torch
    from licos_b200 import ops

    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
    eb = net.entropy_bottleneck
    eb._packed = ops.EbPacked.__new__(ops.EbPacked)           # what packed_params() leaves behind (no GPU here)
    eb._packed.p = __import__("licos_b200")._lib.EbParams()   # a ctypes Structure with pointers: not picklable
    eb._packed_key = ("stale",)
    clone = copy.deepcopy(net)
    assert clone.entropy_bottleneck._packed is None and clone.entropy_bottleneck._packed_key is None
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert back.entropy_bottleneck._packed is None
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), back.state_dict().values()))
