"""Per-operator autograd layer over NHWC activations, for the cascaded trainers' model zoo
(ResDeconv, EDSR, SRDenseNetA/B: SURVEY.md section 8 row a14).

Tensors are contiguous ``(N, H, W, C)`` in the activation dtype (``nn.act_dtype()``: fp32 parity mode or bf16); in bf16 mode
thin ones (fewer than 8 channels) are views of pitch-8 buffers (see ``_new_thin``).
Every operator is one ``torch.autograd.Function`` whose forward and backward call the C-ABI kernels
(``ops.*``); nothing here falls back to torch convolutions.  The networks on the CycleGAN hot path
(``nn.RDDBNetB`` etc.) do NOT use this layer - they run whole-network fused schedules.
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, List, Optional, Tuple

import torch

from . import engine as _engine
from . import ops
from ._lib import ENGINE_SIMT, ENGINE_TC, WL_RSCK, WL_RSKC, WL_TC_DGRAD_S2
from .ops import Slice


# ------------------------------------------------------------------------------------------
# packed-weight cache (keyed by parameter identity; refreshed when the optimizer steps)
# ------------------------------------------------------------------------------------------

class _PackCache:
    def __init__(self):
        self.d: Dict[tuple, tuple] = {}

    def get(self, w: torch.Tensor, kind: tuple, build, layout: int, dtype: torch.dtype) -> torch.Tensor:
        key = (id(w), kind, layout, dtype)
        stamp = (w._version, w.data_ptr())
        hit = self.d.get(key)
        if hit is not None and hit[0]() is w and hit[1] == stamp:
            return hit[2]
        with torch.no_grad():
            packed = ops.pack_weights(build().contiguous(), layout, dtype)
        if len(self.d) > 4096:
            self.d = {k: v for k, v in self.d.items() if v[0]() is not None}
        self.d[key] = (weakref.ref(w), stamp, packed)
        return packed


_packs = _PackCache()


def _aligned(*cs: int) -> bool:
    return all(c % 8 == 0 for c in cs)


# Thin tensors (images, logits: fewer than 8 channels) live in bf16 mode in NHWC buffers whose pixel pitch is padded to 8 channels
# (16 bytes), as in nn._io_buf: TMA tensor maps can then address them and the thin layers (64 -> 3 prediction conv, its data and
# weight gradients) run on the tcgen05 kernels instead of the FFMA ones (cascade step: 5.3 + 1.7 + 2.1 ms -> 0.7 + 0.3 + 0.5 ms).
# Between the functions of this module such a tensor travels as the channel-sliced VIEW buf[..., :c] of its buffer; only the
# first c channels are ever read or written.  Anything that calls .contiguous() on it gets an ordinary dense copy.
_THIN_PITCH = 8


def _new_thin(n, h, w, c, like: torch.Tensor) -> torch.Tensor:
    if like.dtype == torch.bfloat16 and c < _THIN_PITCH and not os.environ.get("SRCGAN_B200_NO_THIN_PITCH"):
        return ops.new_buf(n, h, w, _THIN_PITCH, like.dtype, like.device)[..., :c]
    return ops.new_buf(n, h, w, c, like.dtype, like.device)


def _is_padded_view(t: torch.Tensor) -> bool:
    if t.dim() != 4 or t.shape[3] >= _THIN_PITCH or t.is_contiguous():
        return False
    n, h, w, c = t.shape
    return t.stride() == (h * w * _THIN_PITCH, w * _THIN_PITCH, _THIN_PITCH, 1) and t.storage_offset() % _THIN_PITCH == 0


def _canon(t: torch.Tensor) -> torch.Tensor:
    """dense NHWC tensor or a padded thin view, whichever ``t`` already is; anything else is copied to dense"""
    return t if (t.is_contiguous() or _is_padded_view(t)) else t.contiguous()


def _sl(t: torch.Tensor, c0: int = 0, c: Optional[int] = None) -> Slice:
    """Slice over a tensor accepted by _canon()"""
    if _is_padded_view(t):
        n, h, w, cc = t.shape
        base = t.as_strided((n, h, w, _THIN_PITCH), t.stride(), t.storage_offset())
        return Slice(base, c0, cc - c0 if c is None else c)
    return Slice(t, c0, c)


def _select(cin, cout, k, stride, dtype, ho, wo, x_ld=None, y_ld=None):
    """(engine, layout) for a full-tensor conv: the tcgen05 engine needs 16-byte pixel pitches on both sides."""
    if dtype == torch.bfloat16 and _aligned(cin if x_ld is None else x_ld, cout if y_ld is None else y_ld):
        return _engine.select(cin, cout, k, stride, False, dtype, ho, wo)
    return ENGINE_SIMT, WL_RSCK


def _select_wgrad(cin, cout, k, stride, dtype, ho, wo, x_ld=None, y_ld=None):
    if dtype == torch.bfloat16 and _aligned(cin if x_ld is None else x_ld, cout if y_ld is None else y_ld):
        return _engine.select_wgrad(cin, cout, k, stride, False, dtype, ho, wo)
    return ENGINE_SIMT


def _chunks(cout: int, cin: int, k: int, stride: int, dtype, ho: int, wo: int) -> List[Tuple[int, int]]:
    """Output-channel chunks: the tensor-core kernels take cout in {32,64,128,256}; wider layers (ResDeconv's 512,
    the 4*cout of a 2x2 deconvolution) run as 256-wide slices of the same output buffer."""
    if cout > 256 and cout % 256 == 0 and _select(cin, 256, k, stride, dtype, ho, wo)[0] == ENGINE_TC:
        return [(c0, 256) for c0 in range(0, cout, 256)]
    return [(0, cout)]


def _conv_into(x: torch.Tensor, w: torch.Tensor, kind: str, build, bias, y: torch.Tensor, k: int, stride: int, pad: int,
               act=None) -> None:
    """y = act(conv(x, build()) + bias) where build() is an OIHW weight derived from parameter ``w``."""
    cin, cout = x.shape[3], y.shape[3]
    xs, ys = _sl(x), _sl(y)
    for c0, cc in _chunks(cout, cin, k, stride, x.dtype, y.shape[1], y.shape[2]):
        eng, layout = _select(cin, cc, k, stride, x.dtype, y.shape[1], y.shape[2], xs.ld, ys.ld)
        sub = (lambda c0=c0, cc=cc: build()[c0:c0 + cc]) if (c0, cc) != (0, cout) else build
        pk = _packs.get(w, (kind, c0, cc), sub, layout, x.dtype)
        b = None if bias is None else (bias if (c0, cc) == (0, cout) else bias.detach()[c0:c0 + cc].contiguous())
        ops.conv_fprop(xs, pk, b, Slice(ys.buf, c0, cc), k, stride, pad, act=act, engine=eng)


def _new(n, h, w, c, like: torch.Tensor) -> torch.Tensor:
    return ops.new_buf(n, h, w, c, like.dtype, like.device)


def _masked(gy: torch.Tensor, y: Optional[torch.Tensor], act: Optional[float]) -> torch.Tensor:
    gy = _canon(gy)
    if act is None:
        return gy
    gz = _new_thin(*gy.shape, gy) if _is_padded_view(gy) else torch.empty_like(gy)
    ops.act_backward(_sl(gy), _sl(y), _sl(gz), act)
    return gz


# ------------------------------------------------------------------------------------------
# layout boundary
# ------------------------------------------------------------------------------------------

class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        n, c, h, w = x.shape
        y = _new_thin(n, h, w, c, torch.empty((), dtype=dtype, device=x.device))
        ops.nchw_to_nhwc(x, _sl(y))
        return y

    @staticmethod
    def backward(ctx, gy):
        return ops.nhwc_to_nchw(_sl(_canon(gy))), None


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.dt = x.dtype
        return ops.nhwc_to_nchw(_sl(_canon(x)))

    @staticmethod
    def backward(ctx, gy):
        n, c, h, w = gy.shape
        g = _new_thin(n, h, w, c, torch.empty((), dtype=ctx.dt, device=gy.device))
        ops.nchw_to_nhwc(gy, _sl(g))
        return g


def to_nhwc(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return _ToNHWC.apply(x, dtype)


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    return _ToNCHW.apply(x)


# ------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------

class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act):
        x = _canon(x)
        n, h, wd, cin = x.shape
        cout, _, k, _ = w.shape
        ho, wo = (h + 2 * pad - k) // stride + 1, (wd + 2 * pad - k) // stride + 1
        y = _new_thin(n, ho, wo, cout, x)
        _conv_into(x, w, "f", lambda: w.detach(), b, y, k, stride, pad, act)
        ctx.save_for_backward(x, w, y if act is not None else None)
        ctx.cfg = (stride, pad, act, b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        stride, pad, act, has_b = ctx.cfg
        gz = _masked(gy, y, act)
        n, h, wd, cin = x.shape
        cout, _, k, _ = w.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _new_thin(n, h, wd, cin, x)
            if stride == 1 and k - 1 - pad >= 0:
                _conv_into(gz, w, "t", lambda: w.detach().transpose(0, 1).flip(2, 3), None, dx, k, 1, k - 1 - pad)
            elif _aligned(cin, cout) and _engine.tc_dgrad_s2_supported(cin, cout, k, stride, pad, x.dtype):
                pk = _packs.get(w, ("d2",), lambda: w.detach(), WL_TC_DGRAD_S2, x.dtype)
                ops.conv_dgrad(_sl(gz), pk, _sl(dx), k, stride, pad, engine=ENGINE_TC)
            elif (k == 1 and stride == 2 and pad == 0 and _aligned(cin, cout)
                  and _select(cout, cin, 1, 1, x.dtype, gz.shape[1], gz.shape[2])[0] == ENGINE_TC):
                # 1x1 stride-2 projection (ResDeconv's downsample branches): its data gradient is W^T dY at the even positions and
                # zero elsewhere - a 1x1 stride-1 convolution on the tcgen05 engine into a compact buffer + one strided copy,
                # instead of the gather-form FFMA kernel (0.8-0.9 ms per layer at batch 64)
                tmp = _new(n, gz.shape[1], gz.shape[2], cin, x)
                _conv_into(gz, w, "t1", lambda: w.detach().transpose(0, 1).contiguous(), None, tmp, 1, 1, 0)
                dx.zero_()
                dx[:, ::2, ::2, :].copy_(tmp)
            else:
                pk = _packs.get(w, ("d",), lambda: w.detach(), WL_RSKC, x.dtype)
                ops.conv_dgrad(_sl(gz), pk, _sl(dx), k, stride, pad)
        want_w, want_b = ctx.needs_input_grad[1], has_b and ctx.needs_input_grad[2]
        if want_w or want_b:
            dw = torch.empty(w.shape, dtype=torch.float32, device=x.device) if want_w else None
            db = torch.empty(cout, dtype=torch.float32, device=x.device) if want_b else None
            xs, gs = _sl(x), _sl(gz)
            eng = _select_wgrad(cin, cout, k, stride, x.dtype, gz.shape[1], gz.shape[2], xs.ld, gs.ld)
            ops.conv_wgrad(xs, gs, dw, db, k, stride, pad, engine=eng)
        return dx, dw, db, None, None, None


def conv2d(x, weight, bias=None, stride: int = 1, padding: int = 0, act: Optional[float] = None) -> torch.Tensor:
    """nn.Conv2d (square kernel) with an optional fused LeakyReLU(act) (act = 0.0 is ReLU)."""
    return _Conv2d.apply(x, weight, bias, int(stride), int(padding), act)


# ------------------------------------------------------------------------------------------
# transposed convolution
# ------------------------------------------------------------------------------------------

def _deconv_as_1x1(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d weight (cin, cout, 2, 2) -> OIHW weight of the 1x1 conv to (a, b, cout) channels."""
    return w.detach().permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0], 1, 1)


class _Deconv2x2(torch.autograd.Function):
    """ConvTranspose2d(k=2, s=2, p=0, no bias): non-overlapping, so it is a 1x1 convolution to 4*cout channels
    followed by a depth-to-space permutation (same formulation as nn.RDDBNet)."""

    @staticmethod
    def forward(ctx, x, w, act):
        x = x.contiguous()
        n, h, wd, cin = x.shape
        cout = w.shape[1]
        wide = _new(n, h, wd, 4 * cout, x)
        _conv_into(x, w, "dc_f", lambda: _deconv_as_1x1(w), None, wide, 1, 1, 0, act)
        y = _new(n, 2 * h, 2 * wd, cout, x)
        ops.depth_to_space(Slice(wide), Slice(y))
        ctx.save_for_backward(x, w, wide if act is not None else None)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, wide = ctx.saved_tensors
        n, h, wd, cin = x.shape
        cout = w.shape[1]
        gw = _new(n, h, wd, 4 * cout, x)
        if ctx.act is not None:
            ops.space_to_depth(Slice(gy.contiguous()), Slice(gw), mask=Slice(wide), mask_slope=ctx.act)
        else:
            ops.space_to_depth(Slice(gy.contiguous()), Slice(gw))
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _new(n, h, wd, cin, x)
            _conv_into(gw, w, "dc_t", lambda: _deconv_as_1x1(w).transpose(0, 1), None, dx, 1, 1, 0)
        if ctx.needs_input_grad[1]:
            dw1 = torch.empty((4 * cout, cin, 1, 1), dtype=torch.float32, device=x.device)
            eng = _select_wgrad(cin, 4 * cout, 1, 1, x.dtype, h, wd)
            ops.conv_wgrad(Slice(x), Slice(gw), dw1, None, 1, 1, 0, engine=eng)
            dw = dw1.reshape(2, 2, cout, cin).permute(3, 2, 0, 1).contiguous()
        return dx, dw, None


class _ConvTranspose2d(torch.autograd.Function):
    """General ConvTranspose2d: forward is the data gradient of the convolution it transposes, backward-data is
    that convolution, and the weight gradient is its wgrad with the roles of input and output swapped."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, opad, act):
        x = x.contiguous()
        n, h, wd, cin = x.shape
        _, cout, k, _ = w.shape
        ho, wo = (h - 1) * stride - 2 * pad + k + opad, (wd - 1) * stride - 2 * pad + k + opad
        assert (ho + 2 * pad - k) // stride + 1 == h, "unsupported ConvTranspose2d geometry"
        y = _new(n, ho, wo, cout, x)
        # the transposed conv's weight (cin, cout, k, k) is the OIHW weight of the conv y -> x  (O = cin, I = cout)
        if _aligned(cin, cout) and _engine.tc_dgrad_s2_supported(cout, cin, k, stride, pad, x.dtype):
            pk = _packs.get(w, ("d2",), lambda: w.detach(), WL_TC_DGRAD_S2, x.dtype)
            ops.conv_dgrad(Slice(x), pk, Slice(y), k, stride, pad, engine=ENGINE_TC)
        else:
            pk = _packs.get(w, ("d",), lambda: w.detach(), WL_RSKC, x.dtype)
            ops.conv_dgrad(Slice(x), pk, Slice(y), k, stride, pad)
        if b is not None or act is not None:
            y = _bias_act_(y, b, act)
        ctx.save_for_backward(x, w, y if act is not None else None)
        ctx.cfg = (stride, pad, act, b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        stride, pad, act, has_b = ctx.cfg
        gz = _masked(gy.contiguous(), y, act)
        n, h, wd, cin = x.shape
        _, cout, k, _ = w.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _new(n, h, wd, cin, x)
            _conv_into(gz, w, "f", lambda: w.detach(), None, dx, k, stride, pad)
        want_w, want_b = ctx.needs_input_grad[1], has_b and ctx.needs_input_grad[2]
        if want_w:
            dw = torch.empty(w.shape, dtype=torch.float32, device=x.device)
            eng = _select_wgrad(cout, cin, k, stride, x.dtype, h, wd)
            ops.conv_wgrad(Slice(gz), Slice(x), dw, None, k, stride, pad, engine=eng)
        if want_b:
            db = torch.empty(cout, dtype=torch.float32, device=x.device)
            ops.colsum(Slice(gz), db)
        return dx, dw, db, None, None, None, None


def _bias_act_(y: torch.Tensor, b: Optional[torch.Tensor], act: Optional[float]) -> torch.Tensor:
    """Per-channel bias and (Leaky)ReLU after a gather-form transposed convolution, in place, one small kernel."""
    if b is not None or act is not None:
        ops.bias_act_(Slice(y), b, act)
    return y


def conv_transpose2d(x, weight, bias=None, stride: int = 2, padding: int = 0, output_padding: int = 0,
                     act: Optional[float] = None) -> torch.Tensor:
    k = weight.shape[2]
    if k == 2 and stride == 2 and padding == 0 and output_padding == 0 and bias is None:
        return _Deconv2x2.apply(x, weight, act)
    return _ConvTranspose2d.apply(x, weight, bias, int(stride), int(padding), int(output_padding), act)


# ------------------------------------------------------------------------------------------
# normalisation / permutations / joins
# ------------------------------------------------------------------------------------------

class _GroupNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, groups, eps, act):
        x = x.contiguous()
        y = torch.empty_like(x)
        res = Slice(residual.contiguous()) if residual is not None else None
        mean, rstd = ops.gn_forward(Slice(x), Slice(y), gamma.detach().float(), beta.detach().float(), groups, eps,
                                    residual=res, act=act)
        ctx.save_for_backward(x, gamma, mean, rstd, y if act is not None else None)
        ctx.cfg = (groups, act, residual is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, mean, rstd, y = ctx.saved_tensors
        groups, act, has_res = ctx.cfg
        gz = _masked(gy.contiguous(), y, act)
        dx = torch.empty_like(x)
        dgamma = torch.empty(x.shape[3], dtype=torch.float32, device=x.device)
        dbeta = torch.empty_like(dgamma)
        ops.gn_backward(Slice(gz), Slice(x), Slice(dx), gamma.detach().float(), mean, rstd, groups, dgamma, dbeta)
        return dx, dgamma, dbeta, (gz if has_res else None), None, None, None


def group_norm(x, groups: int, weight, bias, eps: float = 1e-5, residual=None, act: Optional[float] = None):
    """y = act(GroupNorm(x) + residual)."""
    return _GroupNorm.apply(x, weight, bias, residual, int(groups), float(eps), act)


class _PixelShuffle(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, r):
        x = x.contiguous()
        n, h, w, c = x.shape
        y = _new(n, h * r, w * r, c // (r * r), x)
        ops.pixel_shuffle(Slice(x), Slice(y), r)
        ctx.r = r
        return y

    @staticmethod
    def backward(ctx, gy):
        r = ctx.r
        n, H, W, c = gy.shape
        gx = _new(n, H // r, W // r, c * r * r, gy)
        ops.pixel_shuffle(Slice(gy.contiguous()), Slice(gx), r, adjoint=True)
        return gx, None


def pixel_shuffle(x, r: int):
    return _PixelShuffle.apply(x, int(r))


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        y = torch.empty_like(a)
        ops.add(Slice(a), Slice(b), Slice(y))
        return y

    @staticmethod
    def backward(ctx, gy):
        return gy, gy


def add(a, b):
    return _Add.apply(a, b)


class _Cat(torch.autograd.Function):
    """Channel concatenation: a strided copy of each part into its channel window (``srcgan_add`` with a zero
    operand would cost an extra read, so this uses torch's strided copy - no arithmetic involved)."""

    @staticmethod
    def forward(ctx, *parts):
        n, h, w, _ = parts[0].shape
        cs = [p.shape[3] for p in parts]
        y = _new(n, h, w, sum(cs), parts[0])
        c0 = 0
        for p, c in zip(parts, cs):
            y[..., c0:c0 + c].copy_(p)
            c0 += c
        ctx.cs = cs
        return y

    @staticmethod
    def backward(ctx, gy):
        out, c0 = [], 0
        for c in ctx.cs:
            out.append(gy[..., c0:c0 + c].contiguous())
            c0 += c
        return tuple(out)


def cat(parts) -> torch.Tensor:
    return _Cat.apply(*parts)
