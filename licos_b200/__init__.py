"""licos_b200: the LICOS / CompressAI learned-codec hot path on B200 (sm_100a).

Importing the package loads (or builds) the CUDA shared library; there is no CPU fallback."""
from . import _lib  # noqa: F401  (fails loudly if liblicos_b200.so cannot be loaded)
from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional, get_scale_table
from .layers import GDN, FusedSequential, LowerBound, NonNegativeParametrizer, conv, deconv
from .losses import RateDistortionLoss, compute_bpp, compute_msssim, compute_psnr
from .models import (CompressionModel, FactorizedPrior, FactorizedPriorReLU, ScaleHyperprior, image_models,
                     model_architectures)
from .optimizers import net_aux_optimizer
from .graphs import GraphedForward, GraphedTrainStep

__version__ = "0.1.0"


def get_model(model, pretrained, in_channels=3, quality=1):
    """/root/reference/licos/model_utils.py:6-49, run against this package's zoo."""
    import torch.nn as nn

    net = image_models[model](quality=quality, pretrained=pretrained) if model in image_models else None
    if net is None:
        raise ValueError("model: " + model + " not supported for raw data.")
    net.entropy_bottleneck = EntropyBottleneck(channels=net.entropy_bottleneck.channels,
                                               filters=(in_channels, in_channels, 3, 3))
    old = net.g_a[0]
    net.g_a[0] = nn.Conv2d(in_channels=in_channels, out_channels=old.out_channels,
                           kernel_size=(old.weight.shape[2], old.weight.shape[3]), stride=old.stride,
                           padding=old.padding)
    old = net.g_s[6]
    net.g_s[6] = nn.ConvTranspose2d(in_channels=old.in_channels, out_channels=in_channels,
                                    kernel_size=(old.weight.shape[2], old.weight.shape[3]), stride=old.stride,
                                    padding=old.padding, output_padding=old.output_padding)
    return net
