"""The torch custom-op layer (licos_b200/torch_ops.py): every op is registered under torch.ops.licos_b200 with a schema and
a Meta implementation (shape / dtype inference without a device), and has no CPU kernel -- this package has no CPU path."""
import pytest
import torch

import licos_b200  # noqa: F401
from licos_b200 import _lib, torch_ops


def test_ops_are_registered_with_schemas():
    for name in torch_ops.OPS:
        packet = getattr(torch.ops.licos_b200, name)
        assert "licos_b200::" + name in str(packet.default._schema)


def test_meta_kernels_infer_shapes_without_a_device():
    x = torch.empty(2, 3, 64, 96, device="meta")
    w = torch.empty(128 * 128, dtype=torch.bfloat16, device="meta")
    y = torch.ops.licos_b200.conv_forward(x, _lib.CONV_5X5_S2, _lib.EPI_GDN, _lib.LAYOUT_NCHW_F32, _lib.LAYOUT_NHWC_BF16, 3, 128, w,
                                          None, None, None)
    assert y.shape == (2, 32, 48, 128) and y.dtype == torch.bfloat16
    z = torch.ops.licos_b200.conv_forward(y, _lib.DECONV_5X5_S2, _lib.EPI_NONE, _lib.LAYOUT_NHWC_BF16, _lib.LAYOUT_NCHW_U8, 128, 3, w,
                                          None, None, None, 255)
    assert z.shape == (2, 3, 64, 96) and z.dtype == torch.uint8
    lat = torch.empty(2, 192, 4, 6, device="meta")
    packed, med = torch.empty(192, 58, device="meta"), torch.empty(192, device="meta")
    lut = torch.ops.licos_b200.eb_build_lut(packed, med, [1, 3, 3, 3, 3, 1], 0, 1e-9)
    outs = torch.ops.licos_b200.eb_eval_fused(lat, packed, med, [1, 3, 3, 3, 3, 1], 0, 1e-9, lut, True, False, True, True, None)
    assert [tuple(o.shape) for o in outs] == [(2, 192, 4, 6), (2, 192, 4, 6), (0,), (2, 4, 6, 192), (2, 192, 4, 6)]
    assert outs[4].dtype == torch.int16 and outs[3].dtype == torch.bfloat16
    assert torch.ops.licos_b200.nchw_to_nhwc_bf16(lat).shape == (2, 4, 6, 192)


def test_no_cpu_kernels():
    x = torch.zeros(1, 3, 16, 16)
    w = torch.zeros(128 * 128, dtype=torch.bfloat16)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.licos_b200.conv_forward(x, 0, 0, 0, 1, 3, 128, w, None, None, None)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.licos_b200.sum_log(x, torch.zeros(1, dtype=torch.float64))
