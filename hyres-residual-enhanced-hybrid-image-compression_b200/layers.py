"""Parameter holders with the reference's module / attribute names.

These classes exist so that ``state_dict()`` keys, default initialisation and attribute
paths (``model.g_a[3].conv_b[0].conv[2].weight`` ...) are identical to the reference's
(models/checkerboard.py:35-88, models/layers/*.py, compressai GDN /
ResidualBottleneckBlock).  They carry no arithmetic of their own: the compute graph is
compiled from them by ``engine.CodecEngine`` into sm_100a kernel launches.  Calling one of
them directly raises -- there is no PyTorch fallback path.
"""
import torch
import torch.nn as nn

from .entropy import LowerBound


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(
            f"{type(self).__name__} is a parameter holder of the B200 hot path; run it through the model-level API "
            "(forward / compress / decompress), which launches the fused sm_100a kernels")


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def conv1x1(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


def conv3x3(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum=0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        self.register_buffer("pedestal", torch.Tensor([self.reparam_offset ** 2]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal


class GDN(_Holder):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def effective(self):
        """(gamma [C,C,1,1], beta [C]) after the non-negative reparametrisation."""
        C = self.beta.numel()
        return self.gamma_reparam(self.gamma).reshape(C, C, 1, 1), self.beta_reparam(self.beta)


class ResidualBottleneckBlock(_Holder):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        if in_ch != out_ch:
            raise ValueError("the hot path only uses ResidualBottleneckBlock(N, N)")
        mid_ch = min(in_ch, out_ch) // 2
        self.conv1 = conv1x1(in_ch, mid_ch)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(mid_ch, mid_ch)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv3 = conv1x1(mid_ch, out_ch)
        self.skip = nn.Identity()


class AttentionBlock(_Holder):
    def __init__(self, N):
        super().__init__()

        class ResidualUnit(_Holder):
            def __init__(self):
                super().__init__()
                self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2),
                                          nn.ReLU(inplace=True), conv1x1(N // 2, N))
                self.relu = nn.ReLU(inplace=True)

        self.conv_a = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit())
        self.conv_b = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit(), conv1x1(N, N))


class CheckboardMaskedConv2d(nn.Conv2d):
    """5x5 conv whose even-parity taps are masked (12 live taps).  The reference multiplies
    ``weight.data`` by the mask on every call (models/layers/checkerboard.py:47); here the
    dead taps are simply never packed."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.zeros_like(self.weight.data))
        self.mask[:, :, 0::2, 1::2] = 1
        self.mask[:, :, 1::2, 0::2] = 1

    def forward(self, x):
        raise RuntimeError("CheckboardMaskedConv2d is a parameter holder of the B200 hot path")


class Quantizer:
    """models/utils/quantization.py -- kept for attribute parity; the fused quantiser kernels
    (csrc/entropy.cu) implement its three modes."""

    def quantize(self, inputs, quantize_type="noise"):
        if quantize_type == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        if quantize_type == "ste":
            return torch.round(inputs) - inputs.detach() + inputs
        return torch.round(inputs)


class SpatialAttention(_Holder):
    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=(kernel_size - 1) // 2, bias=False)
        self.sigmoid = nn.Sigmoid()


class SEBlock(_Holder):
    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())


def dilated_conv(ch_in, ch_out, dilation):
    return nn.Conv2d(ch_in, ch_out, kernel_size=3, padding=dilation, dilation=dilation, bias=True)


class MultiScaleRefine(_Holder):
    def __init__(self, in_channels=3, mid_channels=64):
        super().__init__()
        if in_channels != 3 or mid_channels != 64:
            raise ValueError("the B200 refine kernels are specialised for in_channels=3, mid_channels=64")
        self.conv_in = nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1)
        self.act_in = nn.PReLU()
        self.se_block = SEBlock(mid_channels, reduction=16)

        def make_block():
            return nn.Sequential(dilated_conv(mid_channels, mid_channels, dilation=1), nn.PReLU(),
                                 dilated_conv(mid_channels, mid_channels, dilation=2), nn.PReLU())

        self.scale1 = make_block()
        self.scale2 = make_block()
        self.scale3 = make_block()
        self.spatial_att = SpatialAttention(kernel_size=7)
        self.fusion = nn.Sequential(nn.Conv2d(mid_channels * 3, mid_channels, kernel_size=1), nn.PReLU(),
                                    nn.Conv2d(mid_channels, in_channels, kernel_size=3, padding=1))
