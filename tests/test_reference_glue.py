"""The reference's own glue files, executed unmodified (tests/ref_shim.py), against the CPU oracle and -- for everything
that needs no kernel -- against the product package.  Build-container tests: they skip where /root/reference is absent
(the GPU box); what travels is the fixture they pin, tests/golden/reference_glue.npz.

  * licos/model_utils.py:6-49 get_model            == oracle.get_model (bit-equal state under one seed) == licos_b200.get_model
  * licos/utils.py:65-73 configure_optimizers      the net / aux split over the product's parameter names
  * licos/train.py:148-212, eval_utils.py:145-210  re-run live and compared with the committed fixture (fixture <-> reference)
  * licos/federation_utils.py:8-85                 file-based merge == licos_b200.federated.merge_pair == oracle rule
  * licos/raw_image_folder.py:183-196              DN scaling == the numpy statement the CUDA kernel is tested against
"""
import os
import sys
import tempfile
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402
import reference_cases as RC  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present on this machine")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_glue.npz")


def _sd_equal(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("model,c,q", [("bmshj2018-factorized", 1, 1), ("bmshj2018-factorized", 3, 1),
                                       ("bmshj2018-factorized", 13, 1), ("bmshj2018-factorized-relu", 3, 2),
                                       ("bmshj2018-hyperprior", 3, 6), ("bmshj2018-hyperprior", 13, 1)])
def test_get_model_reference_file_vs_restatements(model, c, q):
    """model_utils.py:19-45 run from the reference tree builds, under one seed, exactly the state the oracle's and the
    product's restatements of it build: same keys, shapes, dtypes and VALUES (same RNG consumption order)."""
    import licos_b200 as L
    from oracle import compressai_ref as R

    with ref_shim.reference("oracle") as ref:
        torch.manual_seed(3)
        ref_oracle = ref.model_utils.get_model(model, False, c, q)
    with ref_shim.reference("licos_b200") as ref:
        torch.manual_seed(3)
        ref_product = ref.model_utils.get_model(model, False, c, q)
    torch.manual_seed(3)
    own_oracle = R.get_model(model, False, c, q)
    torch.manual_seed(3)
    own_product = L.get_model(model, False, c, q)
    _sd_equal(ref_oracle.state_dict(), own_oracle.state_dict())
    _sd_equal(ref_product.state_dict(), own_product.state_dict())
    _sd_equal(ref_oracle.state_dict(), ref_product.state_dict())
    # the mutated surface (model_utils.py:25-45)
    for net in (ref_oracle, ref_product):
        assert net.entropy_bottleneck.filters == (c, c, 3, 3)
        assert net.g_a[0].in_channels == c and net.g_s[6].out_channels == c
        assert tuple(net.g_a[0].kernel_size) == (5, 5) and tuple(net.g_s[6].output_padding) == (1, 1)
    assert type(ref_product).__module__.startswith("licos_b200")
    counts = {("bmshj2018-factorized", 1, 1): 2980737, ("bmshj2018-factorized", 3, 1): 2998147,
              ("bmshj2018-factorized", 13, 1): 3108237}
    if (model, c, q) in counts:
        assert sum(p.numel() for p in ref_product.parameters()) == counts[(model, c, q)]


def test_get_model_rejects_other_architectures():
    for backend in ("oracle", "licos_b200"):
        with ref_shim.reference(backend) as ref:
            with pytest.raises((ValueError, KeyError)):
                ref.model_utils.get_model("mbt2018-mean", False, 3, 1)
    import licos_b200 as L
    with pytest.raises(ValueError):
        L.get_model("mbt2018-mean", False, 3, 1)


def test_configure_optimizers_split_on_product_model():
    """utils.py:65-73 over the product's module: quantiles -> aux Adam(1e-3), the other 42 tensors -> Adam(1e-4)."""
    with ref_shim.reference("licos_b200") as ref:
        net = ref.model_utils.get_model("bmshj2018-factorized", False, 3, 1)
        opt, aux = ref.utils.configure_optimizers(net, types.SimpleNamespace(learning_rate=1e-4, aux_learning_rate=1e-3))
    assert isinstance(opt, torch.optim.Adam) and isinstance(aux, torch.optim.Adam)
    aux_params = [p for g in aux.param_groups for p in g["params"]]
    assert len(aux_params) == 1 and aux_params[0] is net.entropy_bottleneck.quantiles
    assert sum(len(g["params"]) for g in opt.param_groups) == len(list(net.parameters())) - 1 == 42
    assert opt.param_groups[0]["lr"] == 1e-4 and aux.param_groups[0]["lr"] == 1e-3


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def live():
    """The fixture's contents recomputed now from the reference tree (single-threaded like the generator)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_reference_golden as G

    n = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        return G.generate()
    finally:
        torch.set_num_threads(n)


def test_fixture_is_what_the_reference_code_produces(golden, live):
    assert sorted(golden.files) == sorted(live)
    for k in golden.files:
        a, b = golden[k], live[k]
        assert a.shape == b.shape, k
        if a.dtype.kind in "US":
            assert (a == b).all(), k
        elif a.dtype.kind in "iu":
            assert np.array_equal(a, b), k
        else:
            assert np.allclose(a, b, rtol=1e-5, atol=1e-7), k


def test_reference_eval_semantics(golden):
    """What process_img / compute_bpp guarantee, read off the fixture: the crop at eval_utils.py:207 and the clamp at :205."""
    for name, (model, c, q, shape, seed) in RC.EVAL_CASES.items():
        assert tuple(golden[f"eval/{name}/x_hat_shape"]) == (1, *shape), name
        assert tuple(golden[f"eval/{name}/recon_shape"]) == (shape if c > 1 else shape[1:]), name
        lo = golden[f"eval/{name}/x_hat_lowres"]
        assert lo.min() >= 0 and lo.max() <= 1
        lat = golden[f"eval/{name}/lik_shapes"][0]
        assert tuple(lat[2:]) == (-(-shape[1] // 16), -(-shape[2] // 16)), name  # ceil(H / 16): four ceil(H / 2) convs


def test_raw_band_scaling_matches_the_kernel_contract(golden):
    """raw_image_folder.py:192-196 == the numpy statement tests/test_gpu_entropy.py holds licos_raw_dn_to_unit to."""
    dn = RC.dn_band()
    assert int(golden["raw/dn_max"]) == 4095
    full = (dn.astype(np.float64) / 4095).astype(np.float32)
    ubyte = (np.clip(np.rint(dn.astype(np.float64) / 4095 * 255.0), 0, 255).astype(np.uint8) / 255).astype(np.float32)
    assert np.array_equal(golden["raw/full"][0], full)
    assert np.array_equal(golden["raw/ubyte"][0], ubyte)
    assert np.array_equal(RC.eval_image("split_q1").numpy(), golden["raw/ubyte"])


def _tiny_state(seed, R):
    torch.manual_seed(seed)
    net = R.get_model("bmshj2018-factorized", False, 1, 1)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.01 * torch.randn_like(p))
    return net


def test_federation_file_merge_vs_product_rule(tmp_path, monkeypatch):
    """federation_utils.py:8-85 executed twice (first rank starts the central model, second merges into it) equals
    licos_b200.federated.merge_pair and the oracle's restatement key by key, and the merged state loads back."""
    import licos_b200.federated as F
    from oracle import compressai_ref as R

    monkeypatch.chdir(tmp_path)
    cfg = ref_shim.DotMap(save_path=str(tmp_path / "central"), save_checkpoints_over_time=False)
    with ref_shim.reference("oracle") as ref:
        a, b = _tiny_state(1, R), _tiny_state(2, R)
        a_sd = {k: v.clone() for k, v in a.state_dict().items()}
        b_sd = {k: v.clone() for k, v in b.state_dict().items()}
        loss_a, best_a = torch.tensor(3.0), torch.tensor(2.0)
        ref.federation_utils.update_central_model(0, "cpu", 10, a, loss_a, best_a, 0.0, cfg)
        _sd_equal(a.state_dict(), a_sd)  # first rank: central := local, local unchanged
        loss_b, best_b = torch.tensor(5.0), torch.tensor(1.5)
        ref.federation_utils.update_central_model(1, "cpu", 11, b, loss_b, best_b, 1.0, cfg)
        central = torch.load(str(tmp_path / "central.pth.tar"), weights_only=False)["state_dict"]
    assert not os.path.exists(".mpi_lock")
    merged = F.merge_pair(b_sd, a_sd, float(loss_b), float(best_b))
    oracle_merged = R.federated_average(b_sd, a_sd, float(loss_b), float(best_b))
    w = 1.5 / 6.5
    for k in b_sd:
        assert torch.equal(central[k], b.state_dict()[k]), k           # the rank adopted what it wrote
        assert central[k].dtype == merged[k].dtype, k
        assert torch.allclose(central[k].double(), merged[k].double(), rtol=1e-6, atol=1e-7), k
        assert torch.allclose(central[k].double(), oracle_merged[k].double(), rtol=1e-6, atol=1e-7), k
        if b_sd[k].is_floating_point() and b_sd[k].numel():
            assert torch.allclose(central[k], w * b_sd[k] + (1 - w) * a_sd[k], rtol=1e-6, atol=1e-7), k


def test_test_epoch_runs_reference_loop_on_oracle():
    """train.py:262-303: eval-mode loop with AverageMeter; returns the mean RD loss over the loader."""
    from oracle import compressai_ref as R

    with ref_shim.reference("oracle") as ref:
        net = RC.build_weights(ref.model_utils.get_model, ("bmshj2018-factorized", 3, 1))
        crit = R.RateDistortionLoss(lmbda=1e-2)
        g = torch.Generator().manual_seed(5)
        batches = [torch.rand(2, 3, 64, 64, generator=g) for _ in range(2)]
        avg = ref.train.test_epoch(0, 0, batches, net, crit)
    assert not net.training
    with torch.no_grad():
        want = sum(float(crit(net(d), d)["loss"]) for d in batches) / 2
    assert abs(float(avg) - want) <= 1e-4 * abs(want)
