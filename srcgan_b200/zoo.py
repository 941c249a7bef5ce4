"""The cascaded trainers' remaining generators - ResDeconv, EDSR, SRDenseNetA/B - with the reference's module
trees (state_dict compatible) over the NHWC operator layer in ``functional.py``.

Reference: src/model/resdeconv.py (ResDeconv, BasicBlock), src/model/edsr.py (EDSR, ResnetBlock),
src/model/model.py:616-778 (ConvLayer, DenseLayer, DenseBlock, SRDenseNetA, SRDenseNetB).
Inputs and outputs are NCHW fp32 like the reference modules; inside, activations are NHWC in the activation
dtype and every convolution / normalisation / permutation is a kernel from libsrcgan_b200.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as Fn

RELU = 0.0       # LeakyReLU slope 0


def _act_dtype():
    from .nn import act_dtype
    return act_dtype()


def _kaiming_fanout_(module: nn.Module) -> None:
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


def _conv(m: nn.Conv2d, x, act=None):
    return Fn.conv2d(x, m.weight, m.bias, m.stride[0], m.padding[0], act=act)


def _deconv(m: nn.ConvTranspose2d, x, act=None):
    return Fn.conv_transpose2d(x, m.weight, m.bias, m.stride[0], m.padding[0], m.output_padding[0], act=act)


def _gn(m: nn.GroupNorm, x, residual=None, act=None):
    return Fn.group_norm(x, m.num_groups, m.weight, m.bias, m.eps, residual=residual, act=act)


# ------------------------------------------------------------------------------------------
# ResDeconv
# ------------------------------------------------------------------------------------------

def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def deconv(in_planes, out_planes):
    return nn.ConvTranspose2d(in_planes, out_planes, kernel_size=2, stride=2, padding=0, bias=False, output_padding=0)


class BasicBlock(nn.Module):
    """conv3x3 -> GN -> ReLU -> conv3x3 -> GN -> (+identity | +downsample(x)) -> ReLU; the two tail operations are
    fused into the second GroupNorm's apply pass."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, BN="GN"):
        super().__init__()
        if BN != "GN":
            raise NotImplementedError("ResDeconv is built with GroupNorm (the reference default); BN=%r" % BN)
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.GroupNorm(32, planes)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.GroupNorm(32, planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = _gn(self.bn1, _conv(self.conv1, x), act=RELU)
        identity = x if self.downsample is None else _gn(self.downsample[1], _conv(self.downsample[0], x))
        return _gn(self.bn2, _conv(self.conv2, out), residual=identity, act=RELU)


class ResDeconv(nn.Module):
    def __init__(self, src_ch=1, tar_ch=3, block=BasicBlock, layers=(2, 2, 2, 2), BN="GN"):
        super().__init__()
        if BN != "GN":
            raise NotImplementedError("ResDeconv is built with GroupNorm (the reference default); BN=%r" % BN)
        self.src_ch = src_ch
        if isinstance(tar_ch, list):
            tar_ch = sum(tar_ch)
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.GroupNorm(32, 64)
        self.layer1 = self._make_layer(block, 64, layers[0], 1)
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        self.deconv10 = deconv(512, 256)
        self.inplanes = 256
        self.upRes1 = self._make_layer(block, 256, layers[2], 1)
        self.deconv11 = deconv(256, 128)
        self.inplanes = 128
        self.upRes2 = self._make_layer(block, 128, layers[1], 1)
        self.deconv12 = deconv(128, 64)
        self.inplanes = 64
        self.upRes3 = self._make_layer(block, 64, layers[0], 1)
        self.deconv13 = deconv(64, 64)
        self.pred = nn.Conv2d(64, tar_ch, kernel_size=3, stride=1, padding=1, bias=False)
        _kaiming_fanout_(self)

    def _make_layer(self, block, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       nn.GroupNorm(32, planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        if self.src_ch == 1:
            x = torch.cat([x, x, x], dim=1)
        y = Fn.to_nhwc(x, _act_dtype())
        y = _gn(self.bn1, _conv(self.conv1, y), act=RELU)
        y = self.layer4(self.layer3(self.layer2(self.layer1(y))))
        y = self.upRes1(_deconv(self.deconv10, y))
        y = self.upRes2(_deconv(self.deconv11, y))
        y = self.upRes3(_deconv(self.deconv12, y))
        y = _conv(self.pred, _deconv(self.deconv13, y))
        return Fn.to_nchw(y)


# ------------------------------------------------------------------------------------------
# EDSR
# ------------------------------------------------------------------------------------------

class ResnetBlock(nn.Module):
    def __init__(self, num_channel, kernel=3, stride=1, padding=1):
        super().__init__()
        self.conv1 = nn.Conv2d(num_channel, num_channel, kernel, stride, padding)
        self.conv2 = nn.Conv2d(num_channel, num_channel, kernel, stride, padding)
        self.gn = nn.GroupNorm(32, num_channel)           # one affine pair, applied after both convolutions

    def forward(self, x):
        y = _gn(self.gn, _conv(self.conv1, x), act=0.2)
        return _gn(self.gn, _conv(self.conv2, y), residual=x)


class EDSR(nn.Module):
    def __init__(self, in_ch, ou_ch, upscale_factor=2, base_channel=64, num_residuals=50):
        super().__init__()
        self.input_conv = nn.Conv2d(in_ch, base_channel, kernel_size=3, stride=1, padding=1)
        self.residual_layers = nn.Sequential(*[ResnetBlock(base_channel) for _ in range(num_residuals)])
        self.mid_conv = nn.Conv2d(base_channel, base_channel, kernel_size=3, stride=1, padding=1)
        self.upscale_layers = nn.Sequential(*[deconv(base_channel, base_channel)
                                              for _ in range(int(math.log2(upscale_factor)))])
        self.output_conv = nn.Conv2d(base_channel, ou_ch, kernel_size=3, stride=1, padding=1)
        _kaiming_fanout_(self)

    def forward(self, x):
        y = _conv(self.input_conv, Fn.to_nhwc(x, _act_dtype()))
        res = y
        y = self.residual_layers(y)
        y = Fn.add(_conv(self.mid_conv, y), res)
        for dc in self.upscale_layers:
            y = _deconv(dc, y)
        return Fn.to_nchw(_conv(self.output_conv, y))


# ------------------------------------------------------------------------------------------
# SRDenseNet A / B
# ------------------------------------------------------------------------------------------

class ConvLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=kernel_size // 2)

    def forward(self, x):
        return _conv(self.conv, x, act=RELU)


class DenseLayer(ConvLayer):
    def forward(self, x):
        return Fn.cat([x, _conv(self.conv, x, act=RELU)])


class DenseBlock(nn.Module):
    def __init__(self, in_channels, growth_rate, num_layers):
        super().__init__()
        block = [ConvLayer(in_channels, growth_rate, kernel_size=3)]
        for i in range(num_layers - 1):
            block.append(DenseLayer(growth_rate * (i + 1), growth_rate, kernel_size=3))
        self.block = nn.Sequential(*block)

    def forward(self, x):
        return Fn.cat([x, self.block(x)])


class _SRDenseNet(nn.Module):
    def __init__(self, in_nc, out_nc, nb_channel=1, growth_rate=16, num_blocks=8, num_layers=8, mode="x2"):
        super().__init__()
        self.mode = mode
        self.conv_first = nn.Conv2d(in_nc, 1, 3, 1, 1, bias=True)
        self.conv = ConvLayer(nb_channel, growth_rate * num_layers, 3)
        self.dense_blocks = nn.Sequential(*[DenseBlock(growth_rate * num_layers * (i + 1), growth_rate, num_layers)
                                            for i in range(num_blocks)])
        self.bottleneck = nn.Sequential(
            nn.Conv2d(growth_rate * num_layers + growth_rate * num_layers * num_blocks, 256, kernel_size=1),
            nn.ReLU(inplace=True))
        self.deconv = nn.Sequential(self._resampler(), nn.ReLU(inplace=True))
        self.reconstruction = nn.Conv2d(256, nb_channel, kernel_size=3, padding=1)
        self.conv_last = nn.Conv2d(1, out_nc, 3, 1, 1, bias=True)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight.data, nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias.data)

    def _resampler(self) -> nn.Module:
        raise NotImplementedError

    def forward(self, x):
        y = _conv(self.conv_first, Fn.to_nhwc(x, _act_dtype()))
        y = self.dense_blocks(self.conv(y))
        y = _conv(self.bottleneck[0], y, act=RELU)
        rs = self.deconv[0]
        for _ in range({"x2": 1, "x4": 2}.get(self.mode, 0)):         # the SAME resampler is applied twice for x4
            y = _deconv(rs, y, act=RELU) if isinstance(rs, nn.ConvTranspose2d) else _conv(rs, y, act=RELU)
        return Fn.to_nchw(_conv(self.conv_last, _conv(self.reconstruction, y)))


class SRDenseNetA(_SRDenseNet):
    """Upsampling variant: ConvTranspose2d(256, 256, k3, s2, p1, output_padding 1) + ReLU."""

    def _resampler(self):
        return nn.ConvTranspose2d(256, 256, kernel_size=3, stride=2, padding=1, output_padding=1)


class SRDenseNetB(_SRDenseNet):
    """Downsampling variant: Conv2d(256, 256, k3, s2, p1) + ReLU."""

    def _resampler(self):
        return nn.Conv2d(256, 256, kernel_size=3, stride=2, padding=1)
