"""Thin Python wrappers over the C ABI: torch tensors in, raw device pointers out.

All tensors are torch-owned CUDA memory; calls are asynchronous on torch's current stream.
Activations inside the networks are channel-sliced NHWC buffers (``Slice``)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import (ConvParams, DT_BF16, DT_F32, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC, WL_RSCK, WL_RSKC, WL_TC,  # noqa: F401
                   WL_TC_DGRAD_S2, WL_TC_S2)


def dt_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return DT_F32
    if dtype == torch.bfloat16:
        return DT_BF16
    raise TypeError("srcgan_b200: unsupported activation dtype %s" % dtype)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_current_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  ``torch.cuda.current_stream().cuda_stream`` builds a
    Python Stream object on every call (15 us; at ~2 000 launches per step that was 30 ms of host time per step and left the
    GPU waiting for the host at the start of every step that ends in a host synchronisation)."""
    if _raw_stream is not None:
        return _raw_stream(_current_device())
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    """Kernels launch on the CURRENT device's stream: a tensor that lives on another GPU would be dereferenced by the wrong
    device (or silently through peer access), so that is refused; callers switch with ``torch.cuda.device(t.device)``."""
    if not t.is_cuda:
        raise RuntimeError("srcgan_b200: %s must be a CUDA tensor - this library has no CPU path" % what)
    if t.device.index != _current_device():
        raise RuntimeError("srcgan_b200: %s lives on %s but the current CUDA device is %d - wrap the call in "
                           "torch.cuda.device(tensor.device)" % (what, t.device, torch.cuda.current_device()))


class PlanarBuf:
    """A concat buffer stored as 64-channel GROUPS: physically (groups, N, H, W, 64), logically (N, H, W, groups * 64).
    A slice that stays inside one group is an ordinary (pointer, ld = 64) slice; one that spans groups is described to the
    paired-sweep kernel by ``x_group_stride`` (include/srcgan_b200.h).  Why: a dense block's layers write 32-channel slices;
    in an interleaved 192-channel buffer that is 64 bytes at a 384-byte pitch, which costs 28 % on the 64->32 layer against
    a 128-byte pitch (scripts/exp/layout_probe.py)."""
    GROUP = 64

    def __init__(self, n: int, h: int, w: int, c: int, dtype: torch.dtype, device):
        assert c % self.GROUP == 0
        self.groups = torch.empty((c // self.GROUP, n, h, w, self.GROUP), dtype=dtype, device=device)
        self.shape = (n, h, w, c)
        self.dtype, self.device = dtype, self.groups.device

    def dim(self) -> int:
        return 4

    @property
    def is_cuda(self) -> bool:
        return self.groups.is_cuda

    def element_size(self) -> int:
        return self.groups.element_size()

    @property
    def group_stride(self) -> int:
        return self.groups.stride(0)


class Slice:
    """Channels [c0, c0+c) of an NHWC buffer tensor of shape (N, H, W, Ctot) - or of a ``PlanarBuf``."""
    __slots__ = ("buf", "c0", "c")

    def __init__(self, buf, c0: int = 0, c: Optional[int] = None):
        assert buf.dim() == 4 and (isinstance(buf, PlanarBuf) or buf.is_contiguous())
        self.buf, self.c0 = buf, c0
        self.c = buf.shape[3] - c0 if c is None else c
        assert 0 <= c0 and c0 + self.c <= buf.shape[3]

    @property
    def planar(self) -> bool:
        return isinstance(self.buf, PlanarBuf)

    @property
    def gs(self) -> int:
        """group stride in elements if the slice spans more than one 64-channel group of a planar buffer, else 0"""
        if not self.planar:
            return 0
        g = PlanarBuf.GROUP
        if self.c0 % g:
            assert self.c0 // g == (self.c0 + self.c - 1) // g, "a planar slice that spans groups starts at a group boundary"
            return 0
        return self.buf.group_stride if self.c > g else 0

    @property
    def ptr(self) -> int:
        if self.planar:
            g = PlanarBuf.GROUP
            return self.buf.groups.data_ptr() + ((self.c0 // g) * self.buf.group_stride + self.c0 % g) * self.buf.element_size()
        return self.buf.data_ptr() + self.c0 * self.buf.element_size()

    @property
    def ld(self) -> int:
        return PlanarBuf.GROUP if self.planar else self.buf.shape[3]

    @property
    def n(self) -> int:
        return self.buf.shape[0]

    @property
    def h(self) -> int:
        return self.buf.shape[1]

    @property
    def w(self) -> int:
        return self.buf.shape[2]

    @property
    def npix(self) -> int:
        return self.buf.shape[0] * self.buf.shape[1] * self.buf.shape[2]

    @property
    def dtype(self) -> torch.dtype:
        return self.buf.dtype

    @property
    def device(self):
        return self.buf.device

    def view(self) -> torch.Tensor:
        if self.planar:
            g = PlanarBuf.GROUP
            assert self.c0 // g == (self.c0 + self.c - 1) // g, "view() of a planar slice: one group only"
            return self.buf.groups[self.c0 // g][..., self.c0 % g:self.c0 % g + self.c]
        return self.buf[..., self.c0:self.c0 + self.c]

    def group_slices(self):
        """the slice cut at the 64-channel group boundaries (a non-planar slice is returned whole)"""
        if not self.planar:
            return [self]
        g, out, c = PlanarBuf.GROUP, [], self.c0
        while c < self.c0 + self.c:
            e = min(self.c0 + self.c, (c // g + 1) * g)
            out.append(Slice(self.buf, c, e - c))
            c = e
        return out


def new_buf(n: int, h: int, w: int, c: int, dtype: torch.dtype, device) -> torch.Tensor:
    return torch.empty((n, h, w, c), dtype=dtype, device=device)


def new_concat(n: int, h: int, w: int, c: int, dtype: torch.dtype, device, planar: bool):
    """A dense block's concat buffer: planar 64-channel groups (PlanarBuf) or ordinary interleaved NHWC."""
    return PlanarBuf(n, h, w, c, dtype, device) if planar else new_buf(n, h, w, c, dtype, device)


# ------------------------------------------------------------------------------------------
# workspace (one growing byte buffer per device; safe because every use is stream-ordered)
# ------------------------------------------------------------------------------------------
_workspaces = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    key = (torch.device(device).index, _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------

def pack_weights(w_oihw: torch.Tensor, layout: int, dtype: torch.dtype) -> torch.Tensor:
    """fp32 OIHW -> packed engine layout (a flat tensor of ``dtype``)."""
    _require_cuda(w_oihw, "weight")
    w = w_oihw.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin, kh, kw = w.shape
    lib = _lib.load()
    nbytes = lib.srcgan_packed_weight_bytes(cout, cin, kh, kw, layout, dt_code(dtype))
    out = torch.empty(nbytes // torch.empty((), dtype=dtype).element_size(), dtype=dtype, device=w.device)
    _lib.check(lib.srcgan_pack_weights(w.data_ptr(), cout, cin, kh, kw, layout, dt_code(dtype), out.data_ptr(),
                                       _stream()), "pack_weights")
    return out


# ------------------------------------------------------------------------------------------
# per-launch timing (used by bench.py for the roofline line; off by default)
# ------------------------------------------------------------------------------------------

class KernelTimer:
    """CUDA-event timing of every convolution launch on the launching stream, keyed by kernel family."""

    def __init__(self):
        self.enabled = False
        self.by_shape = False  # key every record by "kernel | kind n x h x w cin->cout k s" (scripts/profile_shapes.py)
        self.records = []      # (key, flops, start_event, end_event, algorithmic bytes)

    def reset(self):
        self.records = []

    def summary(self):
        """-> {key: dict(launches, ms, flops)} ; call after torch.cuda.synchronize()"""
        out = {}
        for key, flops, a, b, nbytes in self.records:
            d = out.setdefault(key, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += a.elapsed_time(b)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


timer = KernelTimer()


class _Timed:
    def __init__(self, kind: str, p: ConvParams, extra: Optional[ConvParams] = None):
        self.on = timer.enabled
        if self.on:
            eng = {ENGINE_SIMT: "simt", ENGINE_TC: "tc", ENGINE_AUTO: "auto"}[p.engine]
            self.key = "conv_%s_%s" % (kind, eng)          # replaced in __exit__ by the kernel the library launched
            self.tc = p.engine == ENGINE_TC and kind != "wgrad"
            if p.engine == ENGINE_TC and kind == "wgrad" and p.kh == 3 and p.stride == 1 and (p.cout == 32 or p.cout % 64 == 0):
                # wgrad = the tcgen05 kernel + a split reduce; name the former (csrc/conv_tc.cu: wgrad_halo_ok)
                stacked = p.pad == 1 and ((p.cout in (32, 64) and p.cin >= 16 and p.cin % 8 == 0) or
                                          (p.cout == 64 and p.cin < 16))                          # wgrad_stack_ok
                self.key = "conv3x3_wgrad_stack_tc" if stacked else "conv3x3_wgrad_halo_tc<%d>" % (64 if p.cout % 64 == 0 else 32)
            self.flops = 2.0 * p.n * p.ho * p.wo * p.cin * p.cout * p.kh * p.kw
            self.shape = "%s %dx%dx%d %d->%d k%d s%d" % (kind, p.n, p.h, p.w, p.cin, p.cout, p.kh, p.stride)
            if extra is not None:
                self.shape += " + %d->%d" % (extra.cin, extra.cout)
            # algorithmic bytes (single read of the input slice, single write / read of the output-side slice; stride and
            # the epilogue's residual / mask operands ignored): what an HBM roofline is computed from
            esz = 2 if p.dtype == DT_BF16 else 4
            self.bytes = float(esz) * p.n * (p.h * p.w * p.cin + p.ho * p.wo * p.cout)
            if extra is not None:          # a fused pair of layers: the second one's FLOPs; its input is already counted
                self.flops += 2.0 * extra.n * extra.ho * extra.wo * extra.cin * extra.cout * extra.kh * extra.kw
                self.bytes += float(esz) * extra.n * extra.ho * extra.wo * extra.cout

    def __enter__(self):
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            if self.tc and exc[0] is None:
                name = _lib.last_kernel()
                if name.startswith("conv"):
                    self.key = name
            if timer.by_shape:
                self.key += " | " + self.shape
            timer.records.append((self.key, self.flops, self.a, self.b, self.bytes))
        return False


# ------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------

def _conv_params(n, h, w, cin, cout, k, stride, pad, upsample, ho, wo, dtype, engine) -> ConvParams:
    p = ConvParams()
    p.n, p.h, p.w, p.cin, p.cout = n, h, w, cin, cout
    p.kh = p.kw = k
    p.stride, p.pad, p.upsample = stride, pad, int(bool(upsample))
    p.ho, p.wo = ho, wo
    p.dtype, p.engine = dt_code(dtype), engine
    p.alpha = 1.0
    return p


def _epilogue(p: ConvParams, bias, act, alpha, r1, beta1, r2, beta2, mask, mask_slope) -> None:
    p.bias = bias.data_ptr() if bias is not None else None
    p.act = 0 if act is None else 1
    p.act_slope = 0.0 if act is None else float(act)
    p.alpha = float(alpha)
    if r1 is not None:
        p.r1, p.r1_ld, p.beta1 = r1.ptr, r1.ld, float(beta1)
    if r2 is not None:
        p.r2, p.r2_ld, p.beta2 = r2.ptr, r2.ld, float(beta2)
    if mask is not None:
        p.mask, p.mask_ld, p.mask_slope = mask.ptr, mask.ld, float(mask_slope)


def conv_fprop(x: Slice, wgt: torch.Tensor, bias: Optional[torch.Tensor], y: Slice, k: int, stride: int = 1,
               pad: int = 1, *, upsample: bool = False, act: Optional[float] = None, alpha: float = 1.0,
               r1: Optional[Slice] = None, beta1: float = 0.0, r2: Optional[Slice] = None, beta2: float = 0.0,
               mask: Optional[Slice] = None, mask_slope: float = 0.0, engine: int = ENGINE_SIMT,
               signbits: Optional[torch.Tensor] = None, maskbits: Optional[torch.Tensor] = None, zero_rows: int = 0,
               reverse: bool = False) -> None:
    """y = epilogue(conv(x, w)); see include/srcgan_b200.h for the epilogue definition.
    ``signbits`` (out) / ``maskbits`` (in): packed LeakyReLU masks, int32 (n, h, w, cout/32) - paired-sweep kernel only.
    ``zero_rows``: separator period of a tall image (rows % zero_rows == 0 are stored as zeros) - paired-sweep kernel only."""
    _require_cuda(x.buf, "conv input")
    p = _conv_params(x.n, x.h, x.w, x.c, y.c, k, stride, pad, upsample, y.h, y.w, x.dtype, engine)
    p.x, p.x_ld, p.wgt, p.y, p.y_ld = x.ptr, x.ld, wgt.data_ptr(), y.ptr, y.ld
    _epilogue(p, bias, act, alpha, r1, beta1, r2, beta2, mask, mask_slope)
    for name, t in (("signbits", signbits), ("maskbits", maskbits)):
        if t is not None:
            assert t.dtype == torch.int32 and t.is_contiguous() and tuple(t.shape) == (y.n, y.h, y.w, y.c // 32), name
            setattr(p, name, t.data_ptr())
    if maskbits is not None:
        p.mask_slope = float(mask_slope)
    p.zero_row_period = int(zero_rows)
    p.flags = 1 if reverse else 0                     # SRCGAN_CONV_FLAG_REVERSE
    p.x_group_stride = x.gs
    assert y.gs == 0 and all(t is None or t.gs == 0 for t in (r1, r2, mask)), "outputs / side operands stay inside one group"
    with _Timed("fprop", p):
        _lib.check(_lib.load().srcgan_conv_fprop(C.byref(p), _stream()), "conv_fprop")


def conv_fprop_pair(xa: Slice, wa: torch.Tensor, ba: Optional[torch.Tensor], ya: Slice,
                    xb: Slice, wb: torch.Tensor, bb: Optional[torch.Tensor], yb: Slice, *, act: Optional[float] = None,
                    signbits=(None, None), maskbits=(None, None), mask_slope: float = 0.0) -> bool:
    """Two consecutive 3x3 layers of a dense block in ONE launch (csrc/conv_pair.cuh): ya = epi(conv(xa, wa)),
    yb = epi(conv(xb, wb)) with xb = [xa | ya] in the same concat buffer.  Returns False WITHOUT launching anything if the
    library cannot fuse this pair (the caller then issues two conv_fprop calls); see srcgan_conv_fprop_pair in the header."""
    _require_cuda(xa.buf, "conv input")
    if os.environ.get("SRCGAN_B200_NO_PAIR_FPROP") or xa.gs or xb.gs or ya.gs or yb.gs:
        return False
    ps = []
    for x, w, b, y, sb, mb in ((xa, wa, ba, ya, signbits[0], maskbits[0]), (xb, wb, bb, yb, signbits[1], maskbits[1])):
        p = _conv_params(x.n, x.h, x.w, x.c, y.c, 3, 1, 1, False, y.h, y.w, x.dtype, ENGINE_TC)
        p.x, p.x_ld, p.wgt, p.y, p.y_ld = x.ptr, x.ld, w.data_ptr(), y.ptr, y.ld
        _epilogue(p, b, act, 1.0, None, 0.0, None, 0.0, None, 0.0)
        for name, t in (("signbits", sb), ("maskbits", mb)):
            if t is not None:
                assert t.dtype == torch.int32 and t.is_contiguous() and tuple(t.shape) == (y.n, y.h, y.w, y.c // 32), name
                setattr(p, name, t.data_ptr())
        if mb is not None:
            p.mask_slope = float(mask_slope)
        ps.append(p)
    lib = _lib.load()
    if not lib.srcgan_conv_fprop_pair_supported(C.byref(ps[0]), C.byref(ps[1])):
        return False
    with _Timed("fprop", ps[0], extra=ps[1]):
        _lib.check(lib.srcgan_conv_fprop_pair(C.byref(ps[0]), C.byref(ps[1]), _stream()), "conv_fprop_pair")
    return True


def conv_dgrad(dy: Slice, wgt_rskc: torch.Tensor, dx: Slice, k: int, stride: int, pad: int, *, alpha: float = 1.0,
               r1: Optional[Slice] = None, beta1: float = 0.0, mask: Optional[Slice] = None,
               mask_slope: float = 0.0, engine: int = ENGINE_SIMT) -> None:
    """dx = epilogue(conv_transpose(dy, w)).  SIMT engine: gather form, any stride, RSKC weights.
    tcgen05 engine: stride 2 (four output-phase convolutions), WL_TC_DGRAD_S2 weights."""
    _require_cuda(dy.buf, "conv grad")
    p = _conv_params(dx.n, dx.h, dx.w, dx.c, dy.c, k, stride, pad, False, dy.h, dy.w, dy.dtype, engine)
    p.x, p.x_ld, p.wgt, p.y, p.y_ld = dy.ptr, dy.ld, wgt_rskc.data_ptr(), dx.ptr, dx.ld
    _epilogue(p, None, None, alpha, r1, beta1, None, 0.0, mask, mask_slope)
    with _Timed("dgrad", p):
        _lib.check(_lib.load().srcgan_conv_dgrad(C.byref(p), _stream()), "conv_dgrad")


def conv_wgrad(x: Slice, dy: Slice, dw: Optional[torch.Tensor], db: Optional[torch.Tensor], k: int, stride: int = 1,
               pad: int = 1, *, upsample: bool = False, accumulate: bool = False, alpha: float = 1.0,
               engine: int = ENGINE_SIMT) -> None:
    """dw (fp32 OIHW) (+)= alpha * sum_pixels x (x) dy ; db (+)= alpha * sum_pixels dy."""
    _require_cuda(x.buf, "conv input")
    assert dy.gs == 0, "dY slices stay inside one group of a planar buffer"
    if x.gs:
        # planar input that spans 64-channel groups: one launch of the kw-stacked kernel per group, each writing its window of
        # input channels of dw (the bias gradient comes out of the first one)
        assert dw is not None and engine == ENGINE_TC and k == 3 and stride == 1 and pad == 1 and not upsample
        conv_wgrad_split(x, dy, (dw, 0, db), None, dy.c, accumulate=accumulate, alpha=alpha)
        return
    if dw is not None:
        assert dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == (dy.c, x.c, k, k), \
            (tuple(dw.shape), (dy.c, x.c, k, k))
    if db is not None:
        assert db.dtype == torch.float32 and db.is_contiguous() and db.numel() == dy.c
    p = _conv_params(x.n, x.h, x.w, x.c, dy.c, k, stride, pad, upsample, dy.h, dy.w, x.dtype, engine)
    p.x, p.x_ld, p.y, p.y_ld = x.ptr, x.ld, dy.ptr, dy.ld
    p.alpha = float(alpha)
    lib = _lib.load()
    nbytes = lib.srcgan_conv_wgrad_workspace_bytes(C.byref(p))
    ws = workspace(nbytes, x.buf.device)
    with _Timed("wgrad", p):
        _lib.check(lib.srcgan_conv_wgrad(C.byref(p), dw.data_ptr() if dw is not None else None,
                                         db.data_ptr() if db is not None else None, int(accumulate),
                                         ws.data_ptr(), ws.numel(), _stream()), "conv_wgrad")


def conv_wgrad_split(x: Slice, dy: Slice, dest0, dest1, split: int, *, accumulate: bool = False, alpha: float = 1.0) -> None:
    """One launch of the kw-stacked tcgen05 wgrad kernel (3x3 stride 1 pad 1), two destinations: output channels [0, split) of
    ``dy`` -> dest0, the rest -> dest1.  dest = (dw, ci0, db) or None: ``dw`` an fp32 OIHW gradient tensor whose input channels
    [ci0, ci0 + x.c) are written, ``db`` the bias gradient of those output channels or None."""
    _require_cuda(x.buf, "conv input")
    assert dy.gs == 0, "dY slices stay inside one group of a planar buffer"
    if x.gs:
        for i, xs in enumerate(x.group_slices()):
            off = xs.c0 - x.c0
            d0 = None if dest0 is None else (dest0[0], dest0[1] + off, dest0[2] if i == 0 else None)
            d1 = None if dest1 is None else (dest1[0], dest1[1] + off, dest1[2] if i == 0 else None)
            conv_wgrad_split(xs, dy, d0, d1, split, accumulate=accumulate, alpha=alpha)
        return
    p = _conv_params(x.n, x.h, x.w, x.c, dy.c, 3, 1, 1, False, dy.h, dy.w, x.dtype, ENGINE_TC)
    p.x, p.x_ld, p.y, p.y_ld = x.ptr, x.ld, dy.ptr, dy.ld
    p.alpha = float(alpha)
    args = []
    for d, co in ((dest0, split), (dest1, dy.c - split)):
        if d is None:
            args += [None, 0, 0, None]
            continue
        dw, ci0, db = d
        assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.shape[0] == co and dw.shape[2:] == (3, 3), tuple(dw.shape)
        assert 0 <= ci0 and ci0 + x.c <= dw.shape[1], (ci0, x.c, tuple(dw.shape))
        assert db is None or (db.dtype == torch.float32 and db.is_contiguous() and db.numel() == co)
        args += [dw.data_ptr(), dw.shape[1], ci0, db.data_ptr() if db is not None else None]
    lib = _lib.load()
    ws = workspace(lib.srcgan_conv_wgrad_workspace_bytes(C.byref(p)), x.buf.device)
    with _Timed("wgrad", p):
        _lib.check(lib.srcgan_conv_wgrad_split(C.byref(p), *args, int(split), int(accumulate), ws.data_ptr(), ws.numel(), _stream()),
                   "conv_wgrad_split")


# ------------------------------------------------------------------------------------------
# glue
# ------------------------------------------------------------------------------------------

def nchw_to_nhwc(src: torch.Tensor, dst: Slice) -> None:
    _require_cuda(src, "input")
    s = src.detach()
    if s.dtype != torch.float32 or not s.is_contiguous():
        s = s.float().contiguous()
    n, c, h, w = s.shape
    assert (dst.n, dst.h, dst.w, dst.c) == (n, h, w, c)
    _lib.check(_lib.load().srcgan_nchw_to_nhwc(s.data_ptr(), n, c, h, w, dst.ptr, dst.ld, dt_code(dst.dtype), _stream()),
               "nchw_to_nhwc")


def nhwc_to_nchw(src: Slice) -> torch.Tensor:
    out = torch.empty((src.n, src.c, src.h, src.w), dtype=torch.float32, device=src.buf.device)
    _lib.check(_lib.load().srcgan_nhwc_to_nchw(src.ptr, src.ld, dt_code(src.dtype), out.data_ptr(), src.n, src.c,
                                               src.h, src.w, _stream()), "nhwc_to_nchw")
    return out


def add(a: Slice, b: Slice, dst: Slice) -> None:
    _lib.check(_lib.load().srcgan_add(a.ptr, a.ld, b.ptr, b.ld, dst.ptr, dst.ld, dst.npix, dst.c, dt_code(dst.dtype),
                                      _stream()), "add")


def bias_act_(y: Slice, bias: Optional[torch.Tensor], act: Optional[float]) -> None:
    """in place: y = leaky(y + bias[channel], act)"""
    b = None
    if bias is not None:
        b = bias.detach()
        if b.dtype != torch.float32 or not b.is_contiguous():
            b = b.float().contiguous()
        assert b.numel() == y.c
    _lib.check(_lib.load().srcgan_bias_act(y.ptr, y.ld, y.npix, y.c, b.data_ptr() if b is not None else None,
                                           int(act is not None), float(act or 0.0), dt_code(y.dtype), _stream()), "bias_act")


def act_backward(dy: Slice, y: Slice, dz: Slice, slope: float) -> None:
    _lib.check(_lib.load().srcgan_act_backward(dy.ptr, dy.ld, y.ptr, y.ld, dz.ptr, dz.ld, dz.npix, dz.c, float(slope),
                                               dt_code(dz.dtype), _stream()), "act_backward")


def colsum(x: Slice, out: torch.Tensor, alpha: float = 1.0, accumulate: bool = False) -> None:
    """out[c] (+)= alpha * sum over all pixels of x[..., c]   (fp32 out, one entry per channel of the slice)"""
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == x.c
    if x.gs:
        for xs in x.group_slices():
            colsum(xs, out[xs.c0 - x.c0:xs.c0 - x.c0 + xs.c], alpha, accumulate)
        return
    lib = _lib.load()
    ws = workspace(lib.srcgan_colsum_workspace_bytes(x.npix, x.c), x.buf.device)
    _lib.check(lib.srcgan_colsum(x.ptr, x.ld, dt_code(x.dtype), x.npix, x.c, out.data_ptr(), float(alpha),
                                 int(accumulate), ws.data_ptr(), ws.numel(), _stream()), "colsum")


def depth_to_space(src: Slice, dst: Slice) -> None:
    assert dst.h == 2 * src.h and dst.w == 2 * src.w and src.c == 4 * dst.c
    _lib.check(_lib.load().srcgan_depth_to_space(src.ptr, src.ld, dst.ptr, dst.ld, src.n, src.h, src.w, dst.c,
                                                 dt_code(dst.dtype), _stream()), "depth_to_space")


def space_to_depth(src: Slice, dst: Slice, mask: Optional[Slice] = None, mask_slope: float = 0.0) -> None:
    assert src.h == 2 * dst.h and src.w == 2 * dst.w and dst.c == 4 * src.c
    _lib.check(_lib.load().srcgan_space_to_depth(
        src.ptr, src.ld, dst.ptr, dst.ld, mask.ptr if mask is not None else None, mask.ld if mask is not None else 0,
        float(mask_slope), dst.n, dst.h, dst.w, src.c, dt_code(dst.dtype), _stream()), "space_to_depth")


def pixel_shuffle(src: Slice, dst: Slice, r: int, adjoint: bool = False) -> None:
    """forward: src (n,h,w,c*r*r) -> dst (n,h*r,w*r,c); adjoint: src (n,h*r,w*r,c) -> dst (n,h,w,c*r*r)"""
    lo, hi = (dst, src) if adjoint else (src, dst)
    assert hi.h == lo.h * r and hi.w == lo.w * r and lo.c == hi.c * r * r
    _lib.check(_lib.load().srcgan_pixel_shuffle(src.ptr, src.ld, dst.ptr, dst.ld, lo.n, lo.h, lo.w, hi.c, r,
                                                dt_code(dst.dtype), int(adjoint), _stream()), "pixel_shuffle")


def upsample2x(src: Slice, dst: Slice) -> None:
    assert dst.h == 2 * src.h and dst.w == 2 * src.w and src.c == dst.c
    _lib.check(_lib.load().srcgan_upsample2x(src.ptr, src.ld, dst.ptr, dst.ld, src.n, src.h, src.w, src.c,
                                             dt_code(dst.dtype), _stream()), "upsample2x")


def upsample2x_adjoint(src: Slice, dst: Slice, mask: Optional[Slice] = None, mask_slope: float = 0.0) -> None:
    assert src.h == 2 * dst.h and src.w == 2 * dst.w and src.c == dst.c
    _lib.check(_lib.load().srcgan_upsample2x_adjoint(
        src.ptr, src.ld, dst.ptr, dst.ld, mask.ptr if mask is not None else None, mask.ld if mask is not None else 0,
        float(mask_slope), dst.n, dst.h, dst.w, dst.c, dt_code(dst.dtype), _stream()), "upsample2x_adjoint")


# ------------------------------------------------------------------------------------------
# batch norm (+ LeakyReLU)
# ------------------------------------------------------------------------------------------

def bn_forward(x: Slice, y: Slice, gamma, beta, running_mean, running_var, training: bool, slope: float,
               momentum: float = 0.1, eps: float = 1e-5):
    c = x.c
    save_mean = torch.empty(c, dtype=torch.float32, device=x.buf.device)
    save_invstd = torch.empty(c, dtype=torch.float32, device=x.buf.device)
    lib = _lib.load()
    ws = workspace(lib.srcgan_bn_workspace_bytes(x.npix, c), x.buf.device)
    _lib.check(lib.srcgan_bn_forward(
        x.ptr, x.ld, y.ptr, y.ld, x.npix, c, dt_code(x.dtype), gamma.data_ptr(), beta.data_ptr(),
        running_mean.data_ptr() if running_mean is not None else None,
        running_var.data_ptr() if running_var is not None else None,
        save_mean.data_ptr(), save_invstd.data_ptr(), int(training), float(momentum), float(eps), float(slope),
        ws.data_ptr(), ws.numel(), _stream()), "bn_forward")
    return save_mean, save_invstd


def bn_backward(dy_post: Slice, y: Slice, x: Slice, dx: Slice, gamma, save_mean, save_invstd, slope: float,
                training: bool, dgamma, dbeta, accumulate: bool = False, beta=None) -> None:
    """``beta`` (the BatchNorm bias): lets the bf16 kernels recompute the LeakyReLU mask from x instead of reading y."""
    lib = _lib.load()
    c = x.c
    ws = workspace(lib.srcgan_bn_workspace_bytes(x.npix, c), x.buf.device)
    _lib.check(lib.srcgan_bn_backward(
        dy_post.ptr, dy_post.ld, y.ptr, y.ld, x.ptr, x.ld, dx.ptr, dx.ld, x.npix, c, dt_code(x.dtype),
        gamma.data_ptr(), beta.data_ptr() if beta is not None else None, save_mean.data_ptr(), save_invstd.data_ptr(),
        float(slope), int(training),
        dgamma.data_ptr() if dgamma is not None else None, dbeta.data_ptr() if dbeta is not None else None,
        int(accumulate), ws.data_ptr(), ws.numel(), _stream()), "bn_backward")


def gn_forward(x: Slice, y: Slice, gamma, beta, groups: int, eps: float = 1e-5, residual: Optional[Slice] = None,
               act: Optional[float] = None):
    n, hw, c = x.n, x.h * x.w, x.c
    mean = torch.empty(n * groups, dtype=torch.float32, device=x.buf.device)
    rstd = torch.empty(n * groups, dtype=torch.float32, device=x.buf.device)
    lib = _lib.load()
    ws = workspace(lib.srcgan_gn_workspace_bytes(n, c), x.buf.device)
    _lib.check(lib.srcgan_gn_forward(x.ptr, x.ld, y.ptr, y.ld, n, hw, c, groups, dt_code(x.dtype), gamma.data_ptr(),
                                     beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), float(eps),
                                     residual.ptr if residual is not None else None,
                                     residual.ld if residual is not None else 0, int(act is not None),
                                     float(act or 0.0), ws.data_ptr(), ws.numel(), _stream()), "gn_forward")
    return mean, rstd


def gn_backward(dy: Slice, x: Slice, dx: Slice, gamma, mean, rstd, groups: int, dgamma, dbeta,
                accumulate: bool = False) -> None:
    n, hw, c = x.n, x.h * x.w, x.c
    lib = _lib.load()
    ws = workspace(lib.srcgan_gn_workspace_bytes(n, c), x.buf.device)
    _lib.check(lib.srcgan_gn_backward(dy.ptr, dy.ld, x.ptr, x.ld, dx.ptr, dx.ld, n, hw, c, groups, dt_code(x.dtype),
                                      gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                      dgamma.data_ptr() if dgamma is not None else None,
                                      dbeta.data_ptr() if dbeta is not None else None, int(accumulate), ws.data_ptr(),
                                      ws.numel(), _stream()), "gn_backward")


# ------------------------------------------------------------------------------------------
# losses / metrics / colour (NCHW fp32 tensors at the module boundary)
# ------------------------------------------------------------------------------------------

def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    _require_cuda(t, what)
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


def loss_fwd_bwd(kind: int, a: torch.Tensor, b, want_grad: bool):
    """Returns (loss 0-dim fp32 tensor, d loss / d a or None).  ``b``: tensor like ``a`` or a python float."""
    a = _f32c(a, "loss input")
    n = a.numel()
    lib = _lib.load()
    ws = workspace(lib.srcgan_loss_workspace_bytes(n), a.device)
    out = torch.empty((), dtype=torch.float32, device=a.device)
    grad = torch.empty_like(a) if want_grad else None
    if isinstance(b, torch.Tensor):
        b = _f32c(b, "loss target")
        assert b.numel() == n, "loss: shape mismatch"
        bp, bs = b.data_ptr(), 0.0
    else:
        bp, bs = None, float(b)
    _lib.check(lib.srcgan_loss_fwd_bwd(kind, a.data_ptr(), bp, bs, n, out.data_ptr(),
                                       grad.data_ptr() if grad is not None else None, ws.data_ptr(), ws.numel(),
                                       _stream()), "loss_fwd_bwd")
    return out, grad


def sq_err_sum(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = _f32c(a, "metric input"), _f32c(b, "metric input")
    lib = _lib.load()
    ws = workspace(lib.srcgan_loss_workspace_bytes(a.numel()), a.device)
    out = torch.empty((), dtype=torch.float32, device=a.device)
    _lib.check(lib.srcgan_metrics_sqerr(a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), ws.data_ptr(),
                                        ws.numel(), _stream()), "metrics_sqerr")
    return out


def angular_error(p: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    p, t = _f32c(p, "metric input"), _f32c(t, "metric input")
    n, c, h, w = p.shape
    out = torch.empty(n, dtype=torch.float32, device=p.device)
    _lib.check(_lib.load().srcgan_metrics_ae(p.data_ptr(), t.data_ptr(), n, c, h, w, out.data_ptr(), _stream()), "ae")
    return out


def minmax(a: torch.Tensor) -> torch.Tensor:
    a = _f32c(a, "metric input")
    out = torch.empty(2, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().srcgan_minmax(a.data_ptr(), a.numel(), out.data_ptr(), _stream()), "minmax")
    return out


def ssim_sums(p: torch.Tensor, t: torch.Tensor, L: float) -> torch.Tensor:
    """Per-image sum of the SSIM map (11x11 'valid' gaussian window)."""
    p, t = _f32c(p, "metric input"), _f32c(t, "metric input")
    n, c, h, w = p.shape
    lib = _lib.load()
    ws = workspace(lib.srcgan_ssim_workspace_bytes(n, c, h, w), p.device)
    out = torch.empty(n, dtype=torch.float32, device=p.device)
    _lib.check(lib.srcgan_ssim(p.data_ptr(), t.data_ptr(), n, c, h, w, float(L), out.data_ptr(), ws.data_ptr(),
                               ws.numel(), _stream()), "ssim")
    return out


def rgb2lab(rgb: torch.Tensor, normalised: bool = True) -> torch.Tensor:
    rgb = _f32c(rgb, "rgb")
    n, c, h, w = rgb.shape
    assert c == 3
    out = torch.empty_like(rgb)
    _lib.check(_lib.load().srcgan_rgb2lab(rgb.data_ptr(), out.data_ptr(), n, h, w, int(normalised), _stream()), "rgb2lab")
    return out


def lab2rgb(lab: torch.Tensor, normalised: bool = True) -> torch.Tensor:
    lab = _f32c(lab, "lab")
    n, c, h, w = lab.shape
    assert c == 3
    out = torch.empty_like(lab)
    _lib.check(_lib.load().srcgan_lab2rgb(lab.data_ptr(), out.data_ptr(), n, h, w, int(normalised), _stream()), "lab2rgb")
    return out


def rgb2lab_u8(img: torch.Tensor) -> torch.Tensor:
    """uint8 (N,H,W,3) image -> normalised LAB (N,3,H,W) float32, float64 arithmetic (dataset.py:148-159)."""
    _require_cuda(img, "image")
    assert img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3
    img = img.contiguous()
    n, h, w, _ = img.shape
    out = torch.empty((n, 3, h, w), dtype=torch.float32, device=img.device)
    _lib.check(_lib.load().srcgan_rgb2lab_u8(img.data_ptr(), out.data_ptr(), n, h, w, _stream()), "rgb2lab_u8")
    return out


def lab2rgb_u8(lab: torch.Tensor) -> torch.Tensor:
    """normalised LAB (N,3,H,W) float32 -> uint8 (N,H,W,3) image with the reference's truncation (dataset.py:94-104)."""
    lab = _f32c(lab, "lab")
    n, c, h, w = lab.shape
    assert c == 3
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=lab.device)
    _lib.check(_lib.load().srcgan_lab2rgb_u8(lab.data_ptr(), out.data_ptr(), n, h, w, _stream()), "lab2rgb_u8")
    return out


_eval_ws = {}


def eval_metrics(pred: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """One kernel launch -> fp32 device tensor of 8 + 5n values: [MSE, PSNR, AE mean, SSIM mean, L, min, max, 0, per-image
    SSIM (n), per-image AE (n), per-image MSE (n), per-image SSIM with its own data range (n), those ranges (n)]
    (see include/srcgan_b200.h); no host synchronisation."""
    p, t = _f32c(pred, "metric input"), _f32c(truth, "metric input")
    assert p.shape == t.shape and p.dim() == 4
    n, c, h, w = p.shape
    lib = _lib.load()
    nbytes = lib.srcgan_eval_metrics_workspace_bytes(n, c, h, w)
    key = (p.device.index, _stream())
    ws = _eval_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=p.device)   # zeroed once: the ticket counter
        _eval_ws[key] = ws
    out = torch.empty(8 + 5 * n, dtype=torch.float32, device=p.device)
    _lib.check(lib.srcgan_eval_metrics(p.data_ptr(), t.data_ptr(), n, c, h, w, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _stream()), "eval_metrics")
    return out


def ssim_backward(pred: torch.Tensor, truth: torch.Tensor, L_dev: torch.Tensor, scale_dev: torch.Tensor,
                  coef: float) -> torch.Tensor:
    """scale_dev * coef * d(sum of the SSIM map)/d pred; L_dev / scale_dev are 1-element fp32 device tensors."""
    p, t = _f32c(pred, "ssim input"), _f32c(truth, "ssim input")
    n, c, h, w = p.shape
    lib = _lib.load()
    ws = workspace(lib.srcgan_ssim_backward_workspace_bytes(n, c, h, w), p.device)
    dp = torch.empty_like(p)
    sc = scale_dev.detach().to(torch.float32).reshape(1).contiguous()
    _lib.check(lib.srcgan_ssim_backward(p.data_ptr(), t.data_ptr(), n, c, h, w, L_dev.data_ptr(), sc.data_ptr(), float(coef),
                                        dp.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "ssim_backward")
    return dp
