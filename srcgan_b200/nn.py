"""B200-native networks of the SRCGAN hot path: the RRDB generators and the patch discriminator.

Each network is ONE ``torch.autograd.Function`` whose forward and backward sequence the C-ABI
kernels of libsrcgan_b200.so over channel-sliced NHWC buffers:

* the dense block's ``torch.cat`` (reference src/model/model.py:205-211) is a 192-channel concat
  buffer that every conv writes its slice of; LeakyReLU, bias, the x0.2 residuals of the RDB
  (:211) and of the RRDB (:233) are conv epilogues;
* the backward of a dense block is run as the *mirrored* dense block over a gradient concat
  buffer [dOut | dZ4 | dZ3 | dZ2 | dZ1] with transposed/flipped weights, so dgrad needs no
  read-modify-write accumulation and uses the same forward conv kernel;
* parameters are ordinary fp32 ``nn.Parameter``s with the reference's names and shapes
  (state_dict compatible); packed engine-layout copies are derived caches keyed on the
  parameter version.

Precision: ``SRCGAN_B200_PRECISION=fp32`` (exact-fp32 parity mode, SIMT engine) or ``bf16``
(default: bf16 activations / weights, fp32 accumulation).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .ops import ENGINE_SIMT, ENGINE_TC, Slice, WL_RSCK, WL_RSKC, WL_TC, WL_TC_DGRAD_S2

LRELU = 0.2

_precision = os.environ.get("SRCGAN_B200_PRECISION", "bf16").lower()


def set_precision(p: str) -> None:
    global _precision
    if p not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _precision = p


def precision() -> str:
    return _precision


def act_dtype() -> torch.dtype:
    return torch.float32 if _precision == "fp32" else torch.bfloat16


def select_engine(cin: int, cout: int, k: int, stride: int, upsample: bool, dtype: torch.dtype, h: int, w: int):
    """(engine, fprop weight layout) for one convolution."""
    from . import engine as _engine  # late import: keeps this module importable on CPU-only boxes
    return _engine.select(cin, cout, k, stride, upsample, dtype, h, w)


def _io_buf(n: int, h: int, w: int, c: int, dt, dev) -> Slice:
    """NHWC buffer for a thin (image / logit) tensor.  In bf16 mode the pixel pitch is padded to 8 channels
    (16 bytes) so that TMA tensor maps can address it; only the first ``c`` channels are ever touched."""
    if dt == torch.bfloat16 and c < 8:
        return Slice(ops.new_buf(n, h, w, 8, dt, dev), 0, c)
    return Slice(ops.new_buf(n, h, w, c, dt, dev))


def _engine_mod():
    from . import engine as _engine
    return _engine


# ------------------------------------------------------------------------------------------
# packed-weight cache
# ------------------------------------------------------------------------------------------

class _PackEntry:
    __slots__ = ("dest", "blocks", "stamp", "params", "cout", "ksize", "k_total")


class _Packs:
    """Packed (engine-layout) copies of the parameters, keyed on ``(param._version, data_ptr)``.  Every in-place update
    torch knows about (optimizer steps, ``load_state_dict``, ``copy_`` under ``no_grad``) bumps the version; a write through
    ``param.data`` does NOT - code that does that must call ``net.invalidate_packs()`` (``dist.broadcast_module_state`` and
    ``checkpoint.load_state`` do).

    Two paths:
    * ``get``     - one tensor at a time: ``build()`` (a torch expression) + ``srcgan_pack_weights``; the SIMT layouts and
                    the few stride-2 / padded tcgen05 layers.
    * ``get_tc``  - stride-1 tcgen05 layout, declared as SOURCE BLOCKS of parameters (transposed / rotated / sliced / scaled,
                    concatenated along K).  All such tensors of a network are re-packed by ONE launch
                    (``srcgan_pack_weights_batch``) the first time one of them is requested after its parameters changed:
                    no ATen flip / cat / mul, ~4 launches per training step instead of ~1100."""

    def __init__(self):
        self.d: Dict[tuple, tuple] = {}
        self.tc: Dict[tuple, _PackEntry] = {}
        self.table = None            # device copy of the block table of every entry in self.tc
        self.table_key = None
        self.repacks = 0             # batched launches so far (tests / bench read it)

    def clear(self) -> None:
        self.d.clear()
        for e in self.tc.values():
            e.stamp = None

    def get(self, key, params, build, layout: int, dtype: torch.dtype) -> torch.Tensor:
        stamp = tuple((p._version, p.data_ptr()) for p in params)
        k = (key, layout, dtype)
        hit = self.d.get(k)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        with torch.no_grad():
            packed = ops.pack_weights(build(), layout, dtype)
        self.d[k] = (stamp, packed)
        return packed

    # ---- batched tcgen05 packs ------------------------------------------------------------------------------------
    @staticmethod
    def _stamp(params):
        return tuple((p._version, p.data_ptr()) for p in params)

    def get_tc(self, key, cout: int, ksize: int, blocks) -> torch.Tensor:
        """``blocks``: [(param OIHW, transposed, n_lo, k_lo, k_len, scale)] concatenated along the GEMM-K axis.
        Not transposed: GEMM-N = the parameter's output channels [n_lo, n_lo + cout), GEMM-K = its input channels
        [k_lo, k_lo + k_len).  Transposed (dgrad): N = input channels, K = output channels, taps rotated by 180 degrees."""
        e = self.tc.get(key)
        if e is None:
            e = _PackEntry()
            e.params = tuple(b[0] for b in blocks)
            e.blocks, e.cout, e.ksize = list(blocks), cout, ksize
            e.k_total = sum(b[4] for b in blocks)
            dev = e.params[0].device
            nbytes = _lib.load().srcgan_packed_weight_bytes(cout, e.k_total, ksize, ksize, WL_TC, _lib.DT_BF16)
            e.dest = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device=dev)       # K / N padding stays zero for good
            e.stamp = None
            self.tc[key] = e
            self.table = None
            self._launch([e])                                   # first use: this tensor alone
            e.stamp = self._stamp(e.params)
            return e.dest
        if e.stamp != self._stamp(e.params):
            if e.dest.device != e.params[0].device:             # the module was moved: start over on the new device
                del self.tc[key]
                self.table = None
                return self.get_tc(key, cout, ksize, blocks)
            self.repack_all()
        return e.dest

    def _table_for(self, entries):
        import ctypes as C
        lib = _lib.load()
        rows, elem0 = [], 0
        slot = (C.c_int32 * 16)()
        ns, bn = C.c_int32(), C.c_int32()
        for e in entries:
            _lib.check(lib.srcgan_pack_slots(e.cout, e.ksize, e.ksize, WL_TC, slot, C.byref(ns), C.byref(bn)), "pack_slots")
            taps = e.ksize * e.ksize
            n_pad = (e.cout + bn.value - 1) // bn.value * bn.value
            nchunks = (e.k_total + 63) // 64
            k0 = 0
            for (p, transposed, n_lo, k_lo, k_len, scale) in e.blocks:
                o, i = p.shape[0], p.shape[1]
                b = _lib.PackBlock()
                b.src, b.out = p.data_ptr(), e.dest.data_ptr()
                if transposed:
                    b.nstride, b.kstride, b.src_off = taps, i * taps, n_lo * taps + k_lo * i * taps
                else:
                    b.nstride, b.kstride, b.src_off = i * taps, taps, n_lo * i * taps + k_lo * taps
                b.elem0, b.n_count, b.k0, b.k_len = elem0, e.cout, k0, k_len
                b.bn, b.nchunks, b.total_slots, b.taps = bn.value, nchunks, ns.value, taps
                b.flip, b.scale = int(bool(transposed)), float(scale)
                for t in range(16):
                    b.slot_off[t] = slot[t]
                rows.append(b)
                elem0 += n_pad * ns.value * k_len
                k0 += k_len
        arr = (_lib.PackBlock * len(rows))(*rows)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        return host.to(entries[0].dest.device), len(rows), elem0

    def _launch(self, entries) -> None:
        with torch.cuda.device(entries[0].dest.device):
            tab, n, total = self._table_for(entries)
            _lib.check(_lib.load().srcgan_pack_weights_batch(tab.data_ptr(), n, total, ops._stream()), "pack_weights_batch")
        self.repacks += 1
        self._keep = tab                                        # alive until the next launch (stream-ordered use)

    def repack_all(self) -> None:
        entries = list(self.tc.values())
        key = tuple(p.data_ptr() for e in entries for p in e.params)
        if self.table is None or self.table_key != key:
            self.table, self.table_n, self.table_total = self._table_for(entries)
            self.table_key = key
        with torch.cuda.device(entries[0].dest.device):
            _lib.check(_lib.load().srcgan_pack_weights_batch(self.table.data_ptr(), self.table_n, self.table_total,
                                                             ops._stream()), "pack_weights_batch")
        self.repacks += 1
        for e in entries:
            e.stamp = self._stamp(e.params)


def _alt_order() -> bool:
    """Every second layer of a dense block walks the images backwards (SRCGAN_CONV_FLAG_REVERSE) so that it starts on what the
    previous layer touched last.  Measured on B200 (scripts/exp/rdb_order.py, profiles/r2_rdb_order.txt): 1.836 vs 1.831 ms per
    dense block, 219.4 vs 218.2 patches/s for the step - the 126 MB L2 does not hold enough of a 1.6 GB concat buffer for
    the order to matter.  Off by default; SRCGAN_B200_ALT_ORDER=1 enables it."""
    return bool(os.environ.get("SRCGAN_B200_ALT_ORDER"))


def _batched_ok(cout: int, ksize: int) -> bool:
    """Shapes the batched packer (csrc/pack_batch.cu) takes: the stride-1 tcgen05 layout's channel counts, <= 16 taps."""
    return (cout in (32, 64, 128, 256) or cout <= 16) and ksize * ksize <= 16 and not os.environ.get("SRCGAN_B200_NO_BATCH_PACK")


def _wT(w: torch.Tensor) -> torch.Tensor:
    """OIHW weight of the stride-1 convolution that computes dgrad: swap in/out, rotate taps 180."""
    return w.detach().transpose(0, 1).flip(2, 3)


class _GradBucket:
    """Flat fp32 buffer that holds the gradients of ALL parameters of one network; the wgrad kernels write straight into
    it and ``param.grad`` becomes a view of it (what DDP calls ``gradient_as_bucket_view``).  Consequences:

    * the data-parallel exchange is ONE ``all_reduce`` over ``flat`` per network - no ``torch.cat``, no copy back;
    * a network that runs several times in one backward pass (G_A / G_B: three times per ``loss_G.backward()``,
      train.py:298-323) accumulates in place: the first node of a pass returns the view to autograd, the later ones add to
      the same memory and return nothing - the ~470 per-parameter ATen adds per step of autograd's input buffer are gone.

    ``seen`` remembers, per parameter, the autograd graph task that wrote it last (``torch._C._current_graph_task_id``)."""
    ALIGN = 64          # floats: every parameter's gradient starts on a 256-byte boundary

    def __init__(self):
        self.flat: Optional[torch.Tensor] = None
        self.off: Dict[int, int] = {}
        self.key = None
        self.seen: Dict[int, int] = {}
        self.writes = 0                 # backward passes that wrote into the bucket (the DP reducer checks it)

    def ensure(self, params, device) -> None:
        key = (tuple(id(p) for p in params), str(device))
        if self.key == key and self.flat is not None:
            return
        off, total = {}, 0
        for p in params:
            if id(p) in off:
                continue
            off[id(p)] = total
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        self.off, self.key, self.seen = off, key, {}

    def view(self, p) -> Optional[torch.Tensor]:
        o = self.off.get(id(p))
        if o is None or p.dtype != torch.float32:
            return None
        return self.flat[o:o + p.numel()].view(p.shape)

    def holds(self, p) -> bool:
        """True if ``p.grad`` currently IS this bucket's view of ``p``."""
        o = self.off.get(id(p))
        g = p.grad
        return (o is not None and g is not None and self.flat is not None and g.dtype == torch.float32
                and g.data_ptr() == self.flat.data_ptr() + 4 * o and g.is_contiguous())


class _GradSink:
    """Collects the parameter gradients produced by the wgrad kernels of ONE backward call of one network (handles shared
    weights: HRconv is applied eight times).  With a bucket the kernels write into the flat buffer; see _GradBucket."""

    def __init__(self, bucket: Optional[_GradBucket] = None, graph_task: int = -1):
        self.g: Dict[int, torch.Tensor] = {}
        self.ret: Dict[int, Optional[torch.Tensor]] = {}
        self.bucket, self.gt = bucket, graph_task

    def slot(self, p: Optional[torch.Tensor], wanted: bool):
        """-> (tensor or None, accumulate)"""
        if p is None or not wanted:
            return None, False
        t = self.g.get(id(p))
        if t is not None:
            return t, True
        v = self.bucket.view(p) if self.bucket is not None else None
        if v is None:
            t = torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
            self.g[id(p)] = self.ret[id(p)] = t
            return t, False
        b = self.bucket
        # accumulate in place when this graph task already wrote the parameter (an earlier node of the same backward
        # pass), or when p.grad still is the bucket view (gradient accumulation over several backward passes)
        inplace = (self.gt != -1 and b.seen.get(id(p)) == self.gt) or b.holds(p)
        b.seen[id(p)] = self.gt
        self.g[id(p)] = v
        self.ret[id(p)] = None if inplace else v
        return v, inplace

    def get(self, p):
        """What the autograd Function returns for ``p``: the gradient tensor, or None if it was added in place."""
        return self.ret.get(id(p))

    def put(self, p, t: torch.Tensor) -> None:
        """Adopt an externally computed gradient (adds if the parameter already has one)."""
        cur = self.g.get(id(p))
        if cur is not None:
            cur.add_(t)
            return
        v, acc = self.slot(p, True)
        if acc:
            v.add_(t)
        else:
            v.copy_(t)


class _NetBase(nn.Module):
    """Shared plumbing: packed weights, conv helpers."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_packs", _Packs())
        object.__setattr__(self, "_bucket", _GradBucket())
        object.__setattr__(self, "_pending_bw", 0)          # forward calls whose backward has not run yet
        object.__setattr__(self, "_grads_ready_hook", None)  # callable(net): every pending backward of this net has run

    def grad_bucket(self) -> "_GradBucket":
        return self.__dict__["_bucket"]

    # nn.Module.__setattr__ would try to register the cache; keep it a plain attribute
    def _pk(self) -> _Packs:
        return self.__dict__["_packs"]

    def invalidate_packs(self) -> None:
        """Drop the packed-weight caches (needed after writes through ``param.data``, which torch does not version)."""
        self._pk().clear()

    def _w_f(self, conv: nn.Conv2d, dtype, layout=WL_RSCK):
        w = conv.weight
        if layout == WL_TC and dtype == torch.bfloat16 and _batched_ok(w.shape[0], w.shape[2]):
            return self._pk().get_tc(("f", id(conv)), w.shape[0], w.shape[2], [(w, False, 0, 0, w.shape[1], 1.0)])
        return self._pk().get(("f", id(conv)), (conv.weight,), lambda: conv.weight.detach(), layout, dtype)

    def _w_t(self, conv: nn.Conv2d, dtype, layout=WL_RSCK):
        w = conv.weight
        if layout == WL_TC and dtype == torch.bfloat16 and _batched_ok(w.shape[1], w.shape[2]):
            return self._pk().get_tc(("t", id(conv)), w.shape[1], w.shape[2], [(w, True, 0, 0, w.shape[0], 1.0)])
        return self._pk().get(("t", id(conv)), (conv.weight,), lambda: _wT(conv.weight).contiguous(), layout, dtype)

    def _w_d(self, conv: nn.Conv2d, dtype):
        return self._pk().get(("d", id(conv)), (conv.weight,), lambda: conv.weight.detach(), WL_RSKC, dtype)

    def _w_d2(self, conv: nn.Conv2d, dtype):
        return self._pk().get(("d2", id(conv)), (conv.weight,), lambda: conv.weight.detach(), WL_TC_DGRAD_S2, dtype)

    def _fprop(self, conv: nn.Conv2d, x: Slice, y: Slice, *, upsample=False, **ep) -> None:
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        eng, layout = select_engine(x.c, y.c, k, s, upsample, x.dtype, y.h, y.w)
        if (eng == ENGINE_TC and k == 3 and s == 1 and p == 1 and not upsample and y.c == 128 and max(y.h, y.w) >= 96):
            # 3x3 layer with 128 output channels on a large map (Decoder conv2, 64 -> 128 at 256^2): two launches of the paired
            # sweep kernel over 64-channel halves instead of the generic implicit GEMM (1.03 -> 0.52 ms at batch 64)
            for h0 in (0, 64):
                wp = self._pk().get(("fhalf", id(conv), h0), (conv.weight,),
                                    lambda h0=h0: conv.weight.detach()[h0:h0 + 64].contiguous(), layout, x.dtype)
                bias = None if conv.bias is None else conv.bias.detach()[h0:h0 + 64].contiguous()
                ep_h = dict(ep)
                for key in ("r1", "r2", "mask"):
                    if ep_h.get(key) is not None:
                        sl = ep_h[key]
                        ep_h[key] = Slice(sl.buf, sl.c0 + h0, 64)
                ops.conv_fprop(x, wp, bias, Slice(y.buf, y.c0 + h0, 64), k, s, p, engine=eng, **ep_h)
            return
        if (eng == ENGINE_SIMT and x.dtype == torch.bfloat16 and y.c <= 4 and k == 4 and s == 1 and not upsample
                and x.c % 64 == 0 and not ep.get("r1") and not ep.get("r2") and not ep.get("mask")
                and select_engine(x.c, 32, k, s, False, x.dtype, y.h, y.w)[0] == ENGINE_TC):
            # patch-logit layer (256 -> 1, 4x4): the one-thread-per-pixel SIMT kernel takes 1.8 ms at batch 64; run it on
            # the tcgen05 implicit GEMM with the output channels zero-padded to 32 and copy the real channels out
            oc = conv.out_channels
            wp = self._pk().get(("f32pad", id(conv)), (conv.weight,),
                                lambda: torch.cat([conv.weight.detach(),
                                                   conv.weight.new_zeros((32 - oc,) + tuple(conv.weight.shape[1:]))]),
                                WL_TC, x.dtype)
            bias = None
            if conv.bias is not None:
                bias = torch.zeros(32, dtype=torch.float32, device=conv.bias.device)
                bias[:oc] = conv.bias.detach()
            tmp = Slice(ops.new_buf(y.n, y.h, y.w, 32, x.dtype, x.buf.device))
            ops.conv_fprop(x, wp, bias, tmp, k, s, p, engine=ENGINE_TC, **ep)
            y.buf[..., y.c0:y.c0 + y.c].copy_(tmp.buf[..., :y.c])
            return
        ops.conv_fprop(x, self._w_f(conv, x.dtype, layout), conv.bias, y, k, s, p, upsample=upsample, engine=eng, **ep)

    def _dgrad(self, conv: nn.Conv2d, dy: Slice, dx: Slice, **ep) -> None:
        """dx = conv^T(dy) with epilogue.  Stride-1 wide convs run as an fprop over transposed weights
        (the path the tensor-core engine shares); strided / thin ones use the gather-form kernel."""
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        thin_tc = s == 1 and _engine_mod().select(dy.c, dx.c, k, 1, False, dy.dtype, dx.h, dx.w)[0] == ENGINE_TC \
            and not (dx.c <= 16 and (ep.get("mask") is not None or ep.get("r1") is not None))
        if s == 1 and ((dy.c > 4 and dx.c > 4) or thin_tc):
            eng, layout = select_engine(dy.c, dx.c, k, 1, False, dy.dtype, dx.h, dx.w)
            ops.conv_fprop(dy, self._w_t(conv, dy.dtype, layout), None, dx, k, 1, k - 1 - p, engine=eng, **ep)
        elif _engine_mod().tc_dgrad_s2_supported(dx.c, dy.c, k, s, p, dy.dtype):
            ops.conv_dgrad(dy, self._w_d2(conv, dy.dtype), dx, k, s, p, engine=ENGINE_TC, **ep)
        elif (dx.c <= 4 and not ep and _engine_mod().tc_dgrad_s2_supported(32, dy.c, k, s, p, dy.dtype)):
            # gradient w.r.t. a 3-channel image through a stride-2 layer (first discriminator layer): the four-phase
            # tcgen05 dgrad with the input channels zero-padded to 32, real channels copied out
            ic = dx.c
            wp = self._pk().get(("d2pad", id(conv)), (conv.weight,),
                                lambda: torch.cat([conv.weight.detach(),
                                                   conv.weight.new_zeros((conv.weight.shape[0], 32 - ic) +
                                                                         tuple(conv.weight.shape[2:]))], dim=1),
                                WL_TC_DGRAD_S2, dy.dtype)
            tmp = Slice(ops.new_buf(dx.n, dx.h, dx.w, 32, dy.dtype, dy.buf.device))
            ops.conv_dgrad(dy, wp, tmp, k, s, p, engine=ENGINE_TC)
            dx.buf[..., dx.c0:dx.c0 + ic].copy_(tmp.buf[..., :ic])
        else:
            ops.conv_dgrad(dy, self._w_d(conv, dy.dtype), dx, k, s, p, **ep)

    def _wgrad(self, conv: nn.Conv2d, x: Slice, dy: Slice, sink: _GradSink, want_w: bool, want_b: bool, *,
               upsample=False, alpha=1.0) -> None:
        dw, acc_w = sink.slot(conv.weight, want_w)
        db, acc_b = sink.slot(conv.bias, want_b)
        if dw is None and db is None:
            return
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        if dw is not None and db is not None and acc_w != acc_b:   # cannot happen (same usage count)
            raise RuntimeError("inconsistent accumulate state")
        from . import engine as _engine
        eng = _engine.select_wgrad(x.c, dy.c, k, s, upsample, x.dtype, dy.h, dy.w)
        if (eng == ENGINE_TC and k == 3 and s == 1 and p == 1 and not upsample and dy.c == 128 and x.c % 8 == 0
                and x.c >= 16 and dw is not None):
            # same layer, weight gradient: two launches of the kw-stacked wgrad kernel over 64-channel halves of dY (any map
            # size - the stacked kernel is tile based; Decoder conv5, 256 -> 128 at 64^2, ran at 161 TFLOP/s on the per-tap one)
            for h0 in (0, 64):
                ops.conv_wgrad(x, Slice(dy.buf, dy.c0 + h0, 64), dw[h0:h0 + 64], None if db is None else db[h0:h0 + 64], k, s, p,
                               accumulate=acc_w or acc_b, alpha=alpha, engine=eng)
            return
        if (eng == ENGINE_SIMT and dw is not None and x.dtype == torch.bfloat16 and dy.c <= 4 and k == 4 and s == 1
                and not upsample and x.c % 64 == 0
                and _engine.select_wgrad(x.c, 64, k, s, False, x.dtype, dy.h, dy.w) == ENGINE_TC):
            # patch-logit layer: weight gradient on the tcgen05 wgrad kernel with dY zero-padded to 64 channels
            oc = dy.c
            dyp = ops.new_buf(dy.n, dy.h, dy.w, 64, x.dtype, x.buf.device)
            dyp.zero_()
            dyp[..., :oc].copy_(dy.buf[..., dy.c0:dy.c0 + oc])
            dwp = torch.empty((64,) + tuple(dw.shape[1:]), dtype=torch.float32, device=dw.device)
            ops.conv_wgrad(x, Slice(dyp), dwp, None, k, s, p, accumulate=False, alpha=alpha, engine=ENGINE_TC)
            if acc_w:
                dw.add_(dwp[:oc])
            else:
                dw.copy_(dwp[:oc])
            if db is not None:
                ops.colsum(dy, db, alpha=alpha, accumulate=acc_b)
            return
        if (eng == ENGINE_SIMT and dw is not None and x.dtype == torch.bfloat16 and x.c <= 4 and k == 4 and s == 2
                and not upsample and dy.c % 64 == 0
                and _engine.select_wgrad(16, dy.c, k, s, False, x.dtype, dy.h, dy.w) == ENGINE_TC):
            # first discriminator layer (3 -> 64, 4x4 stride 2): weight gradient on the tcgen05 wgrad kernel with the image
            # zero-padded to 16 channels
            ic = x.c
            xp = ops.new_buf(x.n, x.h, x.w, 16, x.dtype, x.buf.device)
            xp.zero_()
            xp[..., :ic].copy_(x.buf[..., x.c0:x.c0 + ic])
            dwp = torch.empty((dw.shape[0], 16) + tuple(dw.shape[2:]), dtype=torch.float32, device=dw.device)
            ops.conv_wgrad(Slice(xp), dy, dwp, db, k, s, p, accumulate=False, alpha=alpha, engine=ENGINE_TC) if not acc_b else \
                ops.conv_wgrad(Slice(xp), dy, dwp, None, k, s, p, accumulate=False, alpha=alpha, engine=ENGINE_TC)
            if acc_b and db is not None:
                ops.colsum(dy, db, alpha=alpha, accumulate=True)
            if acc_w:
                dw.add_(dwp[:, :ic])
            else:
                dw.copy_(dwp[:, :ic])
            return
        ops.conv_wgrad(x, dy, dw, db, k, s, p, upsample=upsample, accumulate=acc_w or acc_b, alpha=alpha, engine=eng)


# ------------------------------------------------------------------------------------------
# parameter containers with the reference's module tree (state_dict compatibility)
# ------------------------------------------------------------------------------------------

class ResidualDenseBlock_5(nn.Module):
    """Parameter container of reference model.py:193-202 (conv1..conv5); compute lives in the generator Function."""

    def __init__(self, nf=64, gc=32, bias=True):
        super().__init__()
        self.conv1 = nn.Conv2d(nf, gc, 3, 1, 1, bias=bias)
        self.conv2 = nn.Conv2d(nf + gc, gc, 3, 1, 1, bias=bias)
        self.conv3 = nn.Conv2d(nf + 2 * gc, gc, 3, 1, 1, bias=bias)
        self.conv4 = nn.Conv2d(nf + 3 * gc, gc, 3, 1, 1, bias=bias)
        self.conv5 = nn.Conv2d(nf + 4 * gc, nf, 3, 1, 1, bias=bias)

    def convs(self):
        return [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5]


class RRDB(nn.Module):
    """Parameter container of reference model.py:214-219 (RDB1..3)."""

    def __init__(self, nf, gc=32):
        super().__init__()
        self.RDB1 = ResidualDenseBlock_5(nf, gc)
        self.RDB2 = ResidualDenseBlock_5(nf, gc)
        self.RDB3 = ResidualDenseBlock_5(nf, gc)


def _kaiming_fanout_(module: nn.Module) -> None:
    """Same initialisation pass as the reference constructors (model.py:411-416)."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class Decoder(nn.Module):
    """Parameter container of reference model.py:236-261 (conv1..6 without bias, bn1..6)."""
    PLAN = ((64, 64, 1), (64, 128, 1), (128, 128, 2), (128, 256, 2), (256, 128, 1), (128, 64, 1))
    SLOPE = 0.1

    def __init__(self):
        super().__init__()
        for i, (cin, cout, s) in enumerate(self.PLAN, start=1):
            setattr(self, f"conv{i}", nn.Conv2d(cin, cout, 3, stride=s, padding=1, bias=False))
            setattr(self, f"bn{i}", nn.BatchNorm2d(cout))

    def layers(self):
        return [(getattr(self, f"conv{i}"), getattr(self, f"bn{i}")) for i in range(1, 7)]


# ------------------------------------------------------------------------------------------
# RRDB trunk: forward / backward over concat buffers
# ------------------------------------------------------------------------------------------

class _RRDBGenerator(_NetBase):
    nf: int
    gc: int
    nb: int

    def _rdbs(self) -> List[ResidualDenseBlock_5]:
        out = []
        for rr in self.RRDB_trunk:
            out += [rr.RDB1, rr.RDB2, rr.RDB3]
        return out

    def _dense_wT(self, rdb: ResidualDenseBlock_5, k: int, s5: float, dtype, layout):
        """Packed weights of mirrored-dense-block step k (k=4..1: produces dZ_k; k=0: produces dX)."""
        nf, gc = self.nf, self.gc
        convs = rdb.convs()
        sl = slice(0, nf) if k == 0 else slice(nf + gc * (k - 1), nf + gc * k)

        def build():
            blocks = [_wT(convs[4].weight[:, sl]) * s5]
            for j in range(4, k, -1):                       # conv4 .. conv(k+1)
                blocks.append(_wT(convs[j - 1].weight[:, sl]))
            return torch.cat(blocks, dim=1).contiguous()

        if layout == WL_TC and dtype == torch.bfloat16 and _batched_ok(sl.stop - sl.start, 3):
            # [conv5 | conv4 | ... | conv(k+1)] slices, transposed + rotated, conv5's scaled: blocks of one batched re-pack
            blocks = [(convs[4].weight, True, sl.start, 0, nf, s5)]
            blocks += [(convs[j - 1].weight, True, sl.start, 0, gc, 1.0) for j in range(4, k, -1)]
            return self._pk().get_tc(("dense", id(rdb), k, s5), sl.stop - sl.start, 3, blocks)
        params = tuple(c.weight for c in convs[k:])
        return self._pk().get(("dense", id(rdb), k, s5), params, build, layout, dtype)

    # ---- a chain of dense blocks (groups of three form an RRDB) over concat buffers --------------------
    def _chain_forward(self, rdbs: List[ResidualDenseBlock_5], bufs: List[torch.Tensor], out: Slice,
                       bits: Optional[list] = None, zr: int = 0) -> None:
        """bufs[0][..., :nf] already holds the chain input; the chain output is written to ``out``.
        ``bits`` (a list, filled here): per dense block the packed sign masks of x1..x4 (int32 (n,h,w,1), 4 bytes per pixel)
        when the layer runs on the paired-sweep kernel - the backward steps then read those instead of the 64-byte
        activation slices (dgrad of the gc = 32 layers is HBM-bound: 13-23 % fewer bytes)."""
        nf, gc = self.nf, self.gc
        n, h, w, dt, dev = bufs[0].shape[0], bufs[0].shape[1], bufs[0].shape[2], bufs[0].dtype, bufs[0].device
        alt, flip = _alt_order(), [False]

        def ep_common() -> dict:
            """zero_rows: tall-image mode (see _tall_plan).  reverse: every second launch of the chain walks the images
            backwards, so it starts on what the previous layer touched last - the part the 126 MB L2 still holds."""
            d = {"zero_rows": zr} if zr else {}
            if alt:
                if flip[0]:
                    d["reverse"] = True
                flip[0] = not flip[0]
            return d

        # fused layer pairs (csrc/conv_pair.cuh): bf16, gc = 32, ordinary interleaved buffers, whole lane extent in one CTA pair
        fuse = (dt == torch.bfloat16 and gc == 32 and nf % 64 == 0 and not zr and not alt and not isinstance(bufs[0], ops.PlanarBuf)
                and select_engine(nf, gc, 3, 1, False, dt, h, w)[0] == ENGINE_TC)
        for j, rdb in enumerate(rdbs):
            C = bufs[j]
            convs = rdb.convs()
            row = []
            fused_upto = 0
            for k in range(1, 5):
                sb = None
                if bits is not None and gc == 32 and _engine_mod().sweep_bits_supported(nf + gc * (k - 1), gc, 3, 1, 1, dt, h, w):
                    sb = torch.empty((n, h, w, 1), dtype=torch.int32, device=dev)
                row.append(sb)
            for k in range(1, 5):
                cin_k = nf + gc * (k - 1)
                if k in (1, 3) and fuse:
                    # conv_k and conv_(k+1) in one launch: the prefix is read once, x_k reaches conv_(k+1) through shared memory
                    ca, cb = convs[k - 1], convs[k]
                    if ops.conv_fprop_pair(Slice(C, 0, cin_k), self._w_f(ca, dt, WL_TC), ca.bias, Slice(C, cin_k, gc),
                                           Slice(C, 0, cin_k + gc), self._w_f(cb, dt, WL_TC), cb.bias, Slice(C, cin_k + gc, gc),
                                           act=LRELU, signbits=(row[k - 1], row[k])):
                        fused_upto = k + 1
                if k <= fused_upto:
                    continue
                sb = row[k - 1]
                self._fprop(convs[k - 1], Slice(C, 0, cin_k), Slice(C, cin_k, gc), act=LRELU,
                            **({"signbits": sb} if sb is not None else {}), **ep_common())
            if bits is not None:
                bits.append(row)
            dest = Slice(bufs[j + 1], 0, nf) if j + 1 < len(rdbs) else out
            if j % 3 != 2:      # x5*0.2 + x                                    (model.py:211)
                self._fprop(convs[4], Slice(C), dest, alpha=0.2, r1=Slice(C, 0, nf), beta1=1.0, **ep_common())
            else:               # RDB3: (x5*0.2 + x)*0.2 + rrdb_in              (model.py:211,233)
                self._fprop(convs[4], Slice(C), dest, alpha=0.04, r1=Slice(C, 0, nf), beta1=0.2,
                            r2=Slice(bufs[j - 2], 0, nf), beta2=1.0, **ep_common())

    def _chain_backward(self, rdbs, bufs, Dbuf: List[torch.Tensor], dest: Slice, sink: "_GradSink", W,
                        bits: Optional[list] = None, zr: int = 0) -> None:
        """Dbuf[(len-1) % 4][..., :nf] already holds the gradient w.r.t. the chain output; the gradient
        w.r.t. the chain input (through the chain) is written to ``dest``.  Mirrored dense blocks."""
        nf, gc = self.nf, self.gc
        ctot = nf + 4 * gc
        h, w, dt, dev = bufs[0].shape[1], bufs[0].shape[2], bufs[0].dtype, bufs[0].device
        alt, flip = _alt_order(), [False]
        fuse = (dt == torch.bfloat16 and gc == 32 and nf % 64 == 0 and not zr and not alt and not isinstance(bufs[0], ops.PlanarBuf)
                and not isinstance(Dbuf[0], ops.PlanarBuf))

        def ep_common() -> dict:
            d = {"zero_rows": zr} if zr else {}
            if alt:
                if flip[0]:
                    d["reverse"] = True
                flip[0] = not flip[0]
            return d

        for j in range(len(rdbs) - 1, -1, -1):
            rdb, C, D = rdbs[j], bufs[j], Dbuf[j % 4]
            convs = rdb.convs()
            is_rdb3, is_rdb1 = (j % 3 == 2), (j % 3 == 0)
            s5 = 0.04 if is_rdb3 else 0.2
            fused_downto = 5
            for k in (4, 3, 2, 1):
                cin_v = nf + gc * (4 - k)
                eng, layout = select_engine(cin_v, gc, 3, 1, False, dt, h, w)
                mb = bits[j][k - 1] if bits else None
                if mb is not None and not _engine_mod().sweep_bits_supported(cin_v, gc, 3, 1, 1, dt, h, w):
                    mb = None
                if (k in (4, 2) and fuse and eng == ENGINE_TC and layout == WL_TC and mb is not None and bits[j][k - 2] is not None
                        and _engine_mod().sweep_bits_supported(cin_v + gc, gc, 3, 1, 1, dt, h, w)):
                    # mirrored steps k and k-1 (dZ_k, dZ_(k-1)) in one launch
                    if ops.conv_fprop_pair(Slice(D, 0, cin_v), self._dense_wT(rdb, k, s5, dt, layout), None, Slice(D, cin_v, gc),
                                           Slice(D, 0, cin_v + gc), self._dense_wT(rdb, k - 1, s5, dt, layout), None,
                                           Slice(D, cin_v + gc, gc), maskbits=(mb, bits[j][k - 2]), mask_slope=LRELU):
                        fused_downto = k - 1
                if k >= fused_downto:
                    continue
                if mb is not None:
                    ops.conv_fprop(Slice(D, 0, cin_v), self._dense_wT(rdb, k, s5, dt, layout), None,
                                   Slice(D, nf + gc * (4 - k), gc), 3, 1, 1, maskbits=mb, mask_slope=LRELU, engine=eng, **ep_common())
                else:
                    ops.conv_fprop(Slice(D, 0, cin_v), self._dense_wT(rdb, k, s5, dt, layout), None,
                                   Slice(D, nf + gc * (4 - k), gc), 3, 1, 1,
                                   mask=Slice(C, nf + gc * (k - 1), gc), mask_slope=LRELU, engine=eng, **ep_common())
            dst = Slice(Dbuf[(j - 1) % 4], 0, nf) if j > 0 else dest
            eng, layout = select_engine(ctot, nf, 3, 1, False, dt, h, w)
            ops.conv_fprop(Slice(D), self._dense_wT(rdb, 0, s5, dt, layout), None, dst, 3, 1, 1,
                           r1=Slice(D, 0, nf), beta1=(0.2 if is_rdb3 else 1.0),
                           r2=(Slice(Dbuf[(j + 2) % 4], 0, nf) if is_rdb1 else None), beta2=1.0, engine=eng, **ep_common())
            # bias gradients: the kw-stacked tcgen05 wgrad kernel sums dY out of the slabs it has in shared memory anyway
            # (csrc/conv_tc.cu, tcw4); other engines take one fused column-sum pass over the gradient concat buffer
            from . import engine as _engine
            fused_db = all(_engine.select_wgrad(nf + gc * (k - 1), gc if k < 5 else nf, 3, 1, False, dt, h, w) == ops.ENGINE_TC
                           for k in range(1, 6)) and all(W(c.weight) == W(c.bias) for c in convs)
            paired = (fused_db and dt == torch.bfloat16 and gc == 32 and nf == 64 and not os.environ.get("SRCGAN_B200_NO_WGRAD_PAIRS")
                      and all(W(c.weight) and c.bias is not None and W(c.bias) for c in convs[:4]))
            if paired:
                # conv_k and conv_(k+1) read the same input prefix and their dY slices are neighbours in D ([.. dZ_(k+1) | dZ_k ..]):
                # the common part is ONE 64-output-channel wgrad (N = 192: tensor pipe 95-98 % active, X read once; the
                # 32-channel launches run at N = 96, 62-68 %), plus a 32 -> 32 launch for conv_(k+1)'s own 32 input channels.
                for ka in (1, 3):
                    ca, cb = convs[ka - 1], convs[ka]
                    cin_a, d0 = nf + gc * (ka - 1), nf + gc * (3 - ka)          # common prefix; first channel of dZ_(ka+1) in D
                    dwa, acc_a = sink.slot(ca.weight, True)
                    dba, _ = sink.slot(ca.bias, True)
                    dwb, acc_b = sink.slot(cb.weight, True)
                    dbb, _ = sink.slot(cb.bias, True)
                    if acc_a != acc_b:
                        raise RuntimeError("inconsistent accumulate state of a dense-block layer pair")
                    ops.conv_wgrad_split(Slice(C, 0, cin_a), Slice(D, d0, 2 * gc), (dwb, 0, dbb), (dwa, 0, dba), gc, accumulate=acc_a)
                    ops.conv_wgrad_split(Slice(C, cin_a, gc), Slice(D, d0, gc), (dwb, cin_a, None), None, gc, accumulate=acc_a)
            else:
                for k in range(1, 5):
                    c = convs[k - 1]
                    self._wgrad(c, Slice(C, 0, nf + gc * (k - 1)), Slice(D, nf + gc * (4 - k), gc), sink, W(c.weight),
                                fused_db and W(c.bias))
            self._wgrad(convs[4], Slice(C), Slice(D, 0, nf), sink, W(convs[4].weight), fused_db and W(convs[4].bias), alpha=s5)
            # all five bias gradients = column sums of the gradient concat buffer, one pass
            if not fused_db and any(W(c.bias) for c in convs):
                flat = torch.empty(ctot, dtype=torch.float32, device=dev)
                ops.colsum(Slice(D), flat)
                if W(convs[4].bias):
                    sink.put(convs[4].bias, flat[0:nf].mul_(s5))
                for k in range(1, 5):
                    if W(convs[k - 1].bias):
                        sink.put(convs[k - 1].bias, flat[nf + gc * (4 - k): nf + gc * (5 - k)])

    # ---- "tall image" batching of small maps ---------------------------------------------------------------------------
    # The paired-sweep kernel needs 128 lanes along one image dimension; the G_A trunk runs at 64x64 (BASELINE configs[1]) where
    # only the 4x32-pixel-tile kernels apply (0.23 of the tensor peak).  A batch of n small maps is therefore stacked
    # vertically into ONE image of n*(h+1)+1 rows with an all-zero separator row above, between and below the images: the
    # separator IS the zero padding both neighbours need, so a 3x3 convolution over the tall image equals the per-image
    # convolution as long as every layer writes zeros into the separator rows (conv epilogue option zero_rows = h+1).
    # Weight gradients need nothing: the separators are zero in X and in dY.
    def _tall_plan(self, n: int, h: int, w: int, cin_first: int, dt) -> int:
        """-> separator period (h + 1) if the trunk of this call runs in tall-image mode, else 0"""
        if dt != torch.bfloat16 or max(h, w) >= 96 or os.environ.get("SRCGAN_B200_NO_TALL"):
            return 0
        rows = n * (h + 1) + 1
        if rows < 128:
            return 0
        nf, gc = self.nf, self.gc
        shapes = [(cin_first, nf), (nf, nf), (nf + 4 * gc, nf)] + [(nf + gc * k, gc) for k in range(4)]
        eng = _engine_mod()
        ok = all(eng.sweep_bits_supported(ci, co, 3, 1, 1, dt, rows, w) for ci, co in shapes)
        ok = ok and all(eng.sweep_bits_supported(ci, co, 3, 1, 1, dt, rows, w) for ci, co in
                        [(nf + gc * k, gc) for k in range(4)] + [(nf + 4 * gc, nf)])           # the mirrored backward steps
        return h + 1 if ok else 0

    def _planar_plan(self, h: int, w: int, dt) -> bool:
        """True if the concat buffers of this call are laid out as planar 64-channel groups (ops.PlanarBuf): every layer of the
        dense blocks then writes its 32-channel slice at a 128-byte pitch instead of 384.  Needs every
        dense-block shape on the paired-sweep kernel (the only one that reads across groups)."""
        # OFF by default: measured in the step (profiles/r2_bench_planar_ab.txt) the isolated 28 % on 64->32 shrinks to
        # -1.5 ms for the sweep family while the per-group weight-gradient launches cost +10 ms: 210.4 vs 216.5 patches/s.
        if dt != torch.bfloat16 or self.nf != 64 or self.gc != 32 or not os.environ.get("SRCGAN_B200_PLANAR"):
            return False
        eng = _engine_mod()
        return all(eng.sweep_bits_supported(ci, co, 3, 1, 1, dt, h, w) for ci, co in
                   [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64)])

    @staticmethod
    def _to_tall(src: torch.Tensor, period: int) -> torch.Tensor:
        """(n, h, w, c) -> (1, n*(h+1)+1, w, c) with zero separator rows"""
        n, h, w, c = src.shape
        tall = torch.zeros((1, n * period + 1, w, c), dtype=src.dtype, device=src.device)
        tall[0, 1:].view(n, period, w, c)[:, :h].copy_(src)
        return tall

    @staticmethod
    def _from_tall(tall: torch.Tensor, n: int, h: int) -> torch.Tensor:
        _one, rows, w, c = tall.shape
        return tall[0, 1:].view(n, h + 1, w, c)[:, :h].contiguous()

    def _trunk_forward(self, x_in: Slice, st: dict) -> Slice:
        """conv_first -> nb x RRDB -> trunk_conv (+fea).  Returns fea2; saves buffers in ``st``."""
        nf, gc = self.nf, self.gc
        ctot = nf + 4 * gc
        n0, h0, w0 = x_in.n, x_in.h, x_in.w
        dev, dt = x_in.buf.device, x_in.dtype
        zr = self._tall_plan(n0, h0, w0, x_in.c, dt)
        if zr:
            x_in = Slice(self._to_tall(x_in.buf, zr), x_in.c0, x_in.c)
        zkw = {"zero_rows": zr} if zr else {}
        n, h, w = x_in.n, x_in.h, x_in.w
        rdbs = self._rdbs()
        planar = self._planar_plan(h, w, dt)
        bufs = [ops.new_concat(n, h, w, ctot, dt, dev, planar) for _ in rdbs]
        trunk_out = ops.new_buf(n, h, w, nf, dt, dev)
        fea2 = ops.new_buf(n, h, w, nf, dt, dev)
        self._fprop(self.conv_first, x_in, Slice(bufs[0], 0, nf), **zkw)
        st["bits"] = []
        self._chain_forward(rdbs, bufs, Slice(trunk_out), st["bits"], zr)
        self._fprop(self.trunk_conv, Slice(trunk_out), Slice(fea2), r1=Slice(bufs[0], 0, nf), beta1=1.0, **zkw)
        st["x_in"], st["bufs"], st["trunk_out"], st["fea2"], st["tall"] = x_in, bufs, trunk_out, fea2, (zr, n0, h0)
        st["planar"] = planar
        if zr:
            return Slice(self._from_tall(fea2, n0, h0))
        return Slice(fea2)

    def _trunk_backward(self, st: dict, g_fea2: Slice, sink: _GradSink, want: dict, need_dx: bool) -> Optional[Slice]:
        """Backward of _trunk_forward.  ``want[param_id]`` says which params need grads."""
        nf, gc = self.nf, self.gc
        ctot = nf + 4 * gc
        bufs, trunk_out, x_in = st["bufs"], st["trunk_out"], st["x_in"]
        zr, n0, h0 = st.get("tall", (0, 0, 0))
        zkw = {"zero_rows": zr} if zr else {}
        if zr:
            g_fea2 = Slice(self._to_tall(g_fea2.view().contiguous() if g_fea2.c != g_fea2.ld else g_fea2.buf, zr))
        n, h, w, dev, dt = x_in.n, x_in.h, x_in.w, x_in.buf.device, x_in.dtype
        rdbs = self._rdbs()
        W = lambda p: p is not None and want.get(id(p), False)
        Dbuf = [ops.new_concat(n, h, w, ctot, dt, dev, st.get("planar", False)) for _ in range(4)]
        last = len(rdbs) - 1
        self._wgrad(self.trunk_conv, Slice(trunk_out), g_fea2, sink, W(self.trunk_conv.weight), W(self.trunk_conv.bias))
        self._dgrad(self.trunk_conv, g_fea2, Slice(Dbuf[last % 4], 0, nf), **zkw)
        dfea_trunk = ops.new_buf(n, h, w, nf, dt, dev)
        self._chain_backward(rdbs, bufs, Dbuf, Slice(dfea_trunk), sink, W, st.get("bits"), zr)
        # fea feeds both the trunk and the skip (model.py:421)
        d_fea = ops.new_buf(n, h, w, nf, dt, dev)
        ops.add(Slice(dfea_trunk), g_fea2, Slice(d_fea))
        self._wgrad(self.conv_first, x_in, Slice(d_fea), sink, W(self.conv_first.weight), W(self.conv_first.bias))
        if not need_dx:
            return None
        dx = _io_buf(n, h, w, x_in.c, dt, dev)
        self._dgrad(self.conv_first, Slice(d_fea), dx)          # (the separator rows of dx are never read)
        if zr:
            return Slice(self._from_tall(dx.buf, n0, h0), dx.c0, dx.c)
        return dx


def _flat_params(module: nn.Module) -> List[nn.Parameter]:
    return [p for p in module.parameters()]


def _want_map(ctx_flags, params) -> dict:
    return {id(p): bool(f) for p, f in zip(params, ctx_flags)}


def _device_of(t: torch.Tensor):
    import contextlib
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


class _NetFn(torch.autograd.Function):
    """forward(net, x, *params) -> NCHW fp32; the module implements _forward_impl/_backward_impl."""

    @staticmethod
    def forward(ctx, net, x, *params):
        if not x.is_cuda and not getattr(net, "_host_test_double", False):   # (tests/test_dist_cpu.py drives the bucket /
            raise RuntimeError("srcgan_b200: input must be a CUDA tensor - there is no CPU fallback")   # reducer logic on gloo)
        st: dict = {}
        with _device_of(x):                    # kernels launch on the current device's stream
            out = net._forward_impl(x, st)
        ctx.net, ctx.st, ctx.params = net, st, params
        ctx.counted = any(ctx.needs_input_grad[2:])      # False under no_grad / for frozen parameters: no backward will come
        if ctx.counted:
            net.__dict__["_pending_bw"] += 1
        return out

    @staticmethod
    def backward(ctx, grad_out):
        net, st, params = ctx.net, ctx.st, ctx.params
        flags = ctx.needs_input_grad
        want = _want_map(flags[2:], params)
        bucket = None
        if any(flags[2:]) and not os.environ.get("SRCGAN_B200_NO_GRAD_BUCKET"):
            bucket = net.grad_bucket()
            plist = net.__dict__.get("_bucket_params")
            if plist is None:
                plist = net.__dict__["_bucket_params"] = list(net.parameters())
            bucket.ensure(plist, grad_out.device)
            bucket.writes += 1
        sink = _GradSink(bucket, torch._C._current_graph_task_id())
        with _device_of(grad_out):
            dx = net._backward_impl(st, grad_out, sink, want, flags[1])
        ctx.st = None
        grads = [sink.get(p) if want[id(p)] else None for p in params]
        del sink                                   # no Python reference may outlive this call: autograd adopts the views
        if ctx.counted:
            d = net.__dict__
            d["_pending_bw"] = max(0, d["_pending_bw"] - 1)
            if d["_pending_bw"] == 0 and d["_grads_ready_hook"] is not None and any(flags[2:]):
                d["_grads_ready_hook"](net)
        return (None, dx) + tuple(grads)


# ------------------------------------------------------------------------------------------
# RDDBNetB : the SR generator G_A (reference model.py:396-440)
# ------------------------------------------------------------------------------------------

class RDDBNetB(_RRDBGenerator):
    def __init__(self, in_nc, out_nc, nf, nb=3, gc=32, mode="x2"):
        super().__init__()
        self.nf, self.gc, self.nb, self.mode = nf, gc, nb, mode
        self.conv_first = nn.Conv2d(in_nc, nf, 3, 1, 1, bias=True)
        self.RRDB_trunk = nn.Sequential(*[RRDB(nf, gc) for _ in range(nb)])
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv1 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.HRconv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.conv_last = nn.Conv2d(nf, out_nc, 3, 1, 1, bias=True)
        _kaiming_fanout_(self)

    def _stages(self) -> List[Tuple[nn.Conv2d, bool]]:
        st: List[Tuple[nn.Conv2d, bool]] = []
        if self.mode == "x4":
            st += [(self.upconv1, True), (self.upconv2, True)]
        elif self.mode == "x2":
            st += [(self.upconv1, True), (self.upconv1, False)]      # model.py:429-430
        st += [(self.HRconv, False)] * 8                              # model.py:431-439
        return st

    def forward(self, x):
        return _NetFn.apply(self, x, *_flat_params(self))

    def _forward_impl(self, x: torch.Tensor, st: dict) -> torch.Tensor:
        dt, dev = act_dtype(), x.device
        n, c, h, w = x.shape
        x_in = _io_buf(n, h, w, c, dt, dev)
        ops.nchw_to_nhwc(x, x_in)
        cur = self._trunk_forward(x_in, st)
        acts = [cur]
        abits = [None]      # packed sign masks of the activations (paired-sweep layers): the dgrad steps read 8 B/px, not 128
        ups = {}            # stage index -> materialised nearest-x2 input (bf16 mode: keeps the conv on tcgen05)
        materialise = dt == torch.bfloat16
        for i, (conv, up) in enumerate(self._stages()):
            hh, ww = (cur.h * 2, cur.w * 2) if up else (cur.h, cur.w)
            nxt = Slice(ops.new_buf(n, hh, ww, self.nf, dt, dev))
            sb = None
            if (materialise or not up) and self.nf == 64 and \
                    _engine_mod().sweep_bits_supported(self.nf, self.nf, 3, 1, 1, dt, hh, ww):
                sb = torch.empty((n, hh, ww, 2), dtype=torch.int32, device=dev)
            ep = {"signbits": sb} if sb is not None else {}
            if up and materialise:
                big = Slice(ops.new_buf(n, hh, ww, self.nf, dt, dev))
                ops.upsample2x(cur, big)
                ups[i] = big
                self._fprop(conv, big, nxt, act=LRELU, **ep)
            else:
                self._fprop(conv, cur, nxt, upsample=up, act=LRELU, **ep)
            acts.append(nxt)
            abits.append(sb)
            cur = nxt
        st["ups"], st["abits"] = ups, abits
        out = _io_buf(n, cur.h, cur.w, self.conv_last.out_channels, dt, dev)
        self._fprop(self.conv_last, cur, out)
        st["acts"] = acts
        return ops.nhwc_to_nchw(out)

    def _backward_impl(self, st, grad_out, sink, want, need_dx):
        dt = act_dtype()
        acts = st["acts"]
        dev = grad_out.device
        W = lambda p: p is not None and want.get(id(p), False)
        n = acts[0].n
        last = acts[-1]
        d_o = _io_buf(n, last.h, last.w, self.conv_last.out_channels, dt, dev)
        ops.nchw_to_nhwc(grad_out, d_o)
        self._wgrad(self.conv_last, last, d_o, sink, W(self.conv_last.weight), W(self.conv_last.bias))
        stages = self._stages()
        gz = Slice(ops.new_buf(n, last.h, last.w, self.nf, dt, dev))
        abits = st.get("abits") or [None] * len(acts)

        def mask_ep(i, cin_d, t):
            """epilogue arguments of the dgrad that multiplies by LeakyReLU'(acts[i]); packed bits when both ends are sweep layers"""
            if i <= 0 and t is None:
                return {}
            mb = abits[i] if 0 <= i < len(abits) else None
            if mb is not None and _engine_mod().sweep_bits_supported(cin_d, self.nf, 3, 1, 1, dt, t.h, t.w):
                return {"maskbits": mb, "mask_slope": LRELU}
            return {"mask": t, "mask_slope": LRELU}

        if stages:
            self._dgrad(self.conv_last, d_o, gz, **mask_ep(len(acts) - 1, d_o.c, last))
        else:
            self._dgrad(self.conv_last, d_o, gz)
        for i in range(len(stages) - 1, -1, -1):
            conv, up = stages[i]
            src = acts[i]
            if i in st["ups"]:
                self._wgrad(conv, st["ups"][i], gz, sink, W(conv.weight), W(conv.bias))
            else:
                self._wgrad(conv, src, gz, sink, W(conv.weight), W(conv.bias), upsample=up)
            mask = src if i > 0 else None                    # acts[0] = fea2 is not an activation output
            if up:
                full = Slice(ops.new_buf(n, gz.h, gz.w, self.nf, dt, dev))
                self._dgrad(conv, gz, full)
                nxt = Slice(ops.new_buf(n, src.h, src.w, self.nf, dt, dev))
                ops.upsample2x_adjoint(full, nxt, mask, LRELU)
            else:
                nxt = Slice(ops.new_buf(n, src.h, src.w, self.nf, dt, dev))
                if mask is None:
                    self._dgrad(conv, gz, nxt)
                else:
                    self._dgrad(conv, gz, nxt, **mask_ep(i, gz.c, mask))
            gz = nxt
            if "_debug" in st:
                st["_debug"]["gz%d" % i] = gz.view().float().permute(0, 3, 1, 2).clone()
        dx = self._trunk_backward(st, gz, sink, want, need_dx)
        return ops.nhwc_to_nchw(dx) if dx is not None else None


# ------------------------------------------------------------------------------------------
# RDDBNetA : G_B.  The reference never defines this class; this is the
# documented shim: RDDBNet's constructor (model.py:347-368) + Decoder (model.py:236-289), forward
# = the commented-out one at model.py:370-378.  x4 only.
# ------------------------------------------------------------------------------------------

class RDDBNetA(_RRDBGenerator):
    def __init__(self, in_nc, out_nc, nf, nb, gc=32, mode="x2"):
        super().__init__()
        self.nf, self.gc, self.nb, self.mode = nf, gc, nb, mode
        self.conv_first = nn.Conv2d(in_nc, nf, 3, 1, 1, bias=True)
        self.RRDB_trunk = nn.Sequential(*[RRDB(nf, gc) for _ in range(nb)])
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)      # ctor-compatible, unused by the shim
        self.HRconv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)      # ctor-compatible, unused by the shim
        self.conv_last = nn.Conv2d(nf, out_nc, 3, 1, 1, bias=True)
        _kaiming_fanout_(self)
        self.decode = Decoder()                                   # attached after the init pass, like the shim
        if nf != 64:
            raise ValueError("RDDBNetA shim: Decoder is hard-wired to nf=64 (reference model.py:239)")

    def forward(self, x):
        used = [p for name, p in self.named_parameters() if not name.startswith(("upconv.", "HRconv."))]
        return _NetFn.apply(self, x, *used)

    def _forward_impl(self, x, st):
        dt, dev = act_dtype(), x.device
        n, c, h, w = x.shape
        x_in = _io_buf(n, h, w, c, dt, dev)
        ops.nchw_to_nhwc(x, x_in)
        cur = self._trunk_forward(x_in, st)
        training = self.training
        ys, zs, stats = [], [cur], []
        for conv, bn in self.decode.layers():
            s = conv.stride[0]
            ho, wo = (cur.h + 2 - 3) // s + 1, (cur.w + 2 - 3) // s + 1
            y = Slice(ops.new_buf(n, ho, wo, conv.out_channels, dt, dev))
            self._fprop(conv, cur, y)
            z = Slice(ops.new_buf(n, ho, wo, conv.out_channels, dt, dev))
            use_batch = training or bn.running_mean is None
            sm, si = ops.bn_forward(y, z, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch,
                                    Decoder.SLOPE, bn.momentum if bn.momentum is not None else 0.1, bn.eps)
            if training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
            ys.append(y); zs.append(z); stats.append((sm, si, use_batch))
            cur = z
        out = _io_buf(n, cur.h, cur.w, self.conv_last.out_channels, dt, dev)
        self._fprop(self.conv_last, cur, out)
        st["ys"], st["zs"], st["stats"] = ys, zs, stats
        return ops.nhwc_to_nchw(out)

    def _backward_impl(self, st, grad_out, sink, want, need_dx):
        dt, dev = act_dtype(), grad_out.device
        ys, zs, stats = st["ys"], st["zs"], st["stats"]
        W = lambda p: p is not None and want.get(id(p), False)
        n = zs[0].n
        last = zs[-1]
        d_o = _io_buf(n, last.h, last.w, self.conv_last.out_channels, dt, dev)
        ops.nchw_to_nhwc(grad_out, d_o)
        self._wgrad(self.conv_last, last, d_o, sink, W(self.conv_last.weight), W(self.conv_last.bias))
        g = Slice(ops.new_buf(n, last.h, last.w, last.c, dt, dev))
        self._dgrad(self.conv_last, d_o, g)
        layers = self.decode.layers()
        for i in range(len(layers) - 1, -1, -1):
            conv, bn = layers[i]
            y, z, (sm, si, use_batch) = ys[i], zs[i + 1], stats[i]
            dg, acc_g = sink.slot(bn.weight, W(bn.weight))
            db, acc_b = sink.slot(bn.bias, W(bn.bias))
            ops.bn_backward(g, z, y, g, bn.weight, sm, si, Decoder.SLOPE, use_batch, dg, db, acc_g or acc_b, beta=bn.bias)
            src = zs[i]
            self._wgrad(conv, src, g, sink, W(conv.weight), False)
            nxt = Slice(ops.new_buf(n, src.h, src.w, src.c, dt, dev))
            self._dgrad(conv, g, nxt)
            g = nxt
        dx = self._trunk_backward(st, g, sink, want, need_dx)
        return ops.nhwc_to_nchw(dx) if dx is not None else None


# ------------------------------------------------------------------------------------------
# Cascaded-trainer generators that reuse the dense-block kernels (reference src/model/rddb.py:85-114,
# src/model/srdn.py:53-78); constructed by ``eval(opt.SRModel)(1, 1, opt.up)`` (src/trainCas.py:30)
# ------------------------------------------------------------------------------------------

class RDDBNet(_RRDBGenerator):
    """pkg RDDBNet: conv_first -> nb RRDB -> trunk_conv (+fea) -> log2(up) x [ConvTranspose2d(k2,s2,no bias)
    -> LeakyReLU(0.2)] -> conv_last (no bias).  The non-overlapping deconvolution is a 1x1 convolution to
    4*nf channels (LeakyReLU fused) followed by a depth-to-space permutation."""

    def __init__(self, in_ch, ou_ch, upscale_factor, nf=64, nb=3, gc=32):
        super().__init__()
        import math
        self.nf, self.gc, self.nb, self.upscale_factor = nf, gc, nb, upscale_factor
        self.conv_first = nn.Conv2d(in_ch, nf, 3, 1, 1, bias=True)
        self.RRDB_trunk = nn.Sequential(*[RRDB(nf, gc) for _ in range(nb)])
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        ups: List[nn.Module] = []
        for _ in range(int(math.log2(upscale_factor))):
            ups += [nn.ConvTranspose2d(nf, nf, kernel_size=2, stride=2, padding=0, bias=False, output_padding=0),
                    nn.LeakyReLU(negative_slope=0.2, inplace=True)]
        self.upscale_layers = nn.Sequential(*ups)
        self.conv_last = nn.Conv2d(nf, ou_ch, 3, 1, 1, bias=False)
        _kaiming_fanout_(self)                     # touches nn.Conv2d only, as in the reference (rddb.py:100-105)

    def _deconvs(self) -> List[nn.ConvTranspose2d]:
        return [m for m in self.upscale_layers if isinstance(m, nn.ConvTranspose2d)]

    def _w_deconv(self, dc: nn.ConvTranspose2d, kind: str, dtype, layout):
        """ConvTranspose2d weight (cin, cout, 2, 2) as the OIHW weight of a 1x1 conv to (a,b,cout) channels
        ('f'), or of its transpose ('t': (a,b,cout) -> cin)."""
        def build():
            w = dc.weight.detach()                                      # [ci, co, a, b]
            w1 = w.permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0], 1, 1)   # [(a,b,co), ci, 1, 1]
            return (w1 if kind == "f" else w1.transpose(0, 1)).contiguous()
        return self._pk().get(("deconv", kind, id(dc)), (dc.weight,), build, layout, dtype)

    def forward(self, x):
        return _NetFn.apply(self, x, *_flat_params(self))

    def _forward_impl(self, x, st):
        dt, dev = act_dtype(), x.device
        n, c, h, w = x.shape
        x_in = _io_buf(n, h, w, c, dt, dev)
        ops.nchw_to_nhwc(x, x_in)
        cur = self._trunk_forward(x_in, st)
        nf = self.nf
        ups = []
        for dc in self._deconvs():
            wide = Slice(ops.new_buf(n, cur.h, cur.w, 4 * nf, dt, dev))
            eng, layout = select_engine(nf, 4 * nf, 1, 1, False, dt, cur.h, cur.w)
            ops.conv_fprop(cur, self._w_deconv(dc, "f", dt, layout), None, wide, 1, 1, 0, act=LRELU, engine=eng)
            big = Slice(ops.new_buf(n, 2 * cur.h, 2 * cur.w, nf, dt, dev))
            ops.depth_to_space(wide, big)
            ups.append((cur, wide))
            cur = big
        out = _io_buf(n, cur.h, cur.w, self.conv_last.out_channels, dt, dev)
        self._fprop(self.conv_last, cur, out)
        st["ups"], st["last_in"] = ups, cur
        return ops.nhwc_to_nchw(out)

    def _backward_impl(self, st, grad_out, sink, want, need_dx):
        dt, dev = act_dtype(), grad_out.device
        W = lambda p: p is not None and want.get(id(p), False)
        last = st["last_in"]
        n, nf = last.n, self.nf
        d_o = _io_buf(n, last.h, last.w, self.conv_last.out_channels, dt, dev)
        ops.nchw_to_nhwc(grad_out, d_o)
        self._wgrad(self.conv_last, last, d_o, sink, W(self.conv_last.weight), False)
        g = Slice(ops.new_buf(n, last.h, last.w, nf, dt, dev))
        self._dgrad(self.conv_last, d_o, g)
        deconvs = self._deconvs()
        for i in range(len(deconvs) - 1, -1, -1):
            dc = deconvs[i]
            src, wide = st["ups"][i]
            gw = Slice(ops.new_buf(n, src.h, src.w, 4 * nf, dt, dev))
            ops.space_to_depth(g, gw, mask=wide, mask_slope=LRELU)        # dZ of the 1x1 conv, LeakyReLU mask fused
            if W(dc.weight):
                dw1 = torch.empty((4 * nf, nf, 1, 1), dtype=torch.float32, device=dev)
                eng = _engine_mod().select_wgrad(nf, 4 * nf, 1, 1, False, dt, src.h, src.w)
                ops.conv_wgrad(src, gw, dw1, None, 1, 1, 0, engine=eng)
                sink.put(dc.weight, dw1.reshape(2, 2, nf, nf).permute(3, 2, 0, 1).contiguous())
            nxt = Slice(ops.new_buf(n, src.h, src.w, nf, dt, dev))
            eng, layout = select_engine(4 * nf, nf, 1, 1, False, dt, src.h, src.w)
            ops.conv_fprop(gw, self._w_deconv(dc, "t", dt, layout), None, nxt, 1, 1, 0, engine=eng)
            g = nxt
        dx = self._trunk_backward(st, g, sink, want, need_dx)
        return ops.nhwc_to_nchw(dx) if dx is not None else None


class SRDN(_RRDBGenerator):
    """SRDN: conv_first -> encoder (nb RRDB) -> +fea -> decoder (nb RRDB) -> +fea -> conv_last (no bias).
    No upsampling; ``trunk_conv`` exists (state_dict compatible) but is unused, as in srdn.py:60,69-76."""

    def __init__(self, in_ch, ou_ch, upscale_factor, nf=64, nb=3, gc=32):
        super().__init__()
        self.nf, self.gc, self.nb, self.upscale_factor = nf, gc, nb, upscale_factor
        self.conv_first = nn.Conv2d(in_ch, nf, 3, 1, 1, bias=True)
        self.RRDB_encoder = nn.Sequential(*[RRDB(nf, gc) for _ in range(nb)])
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.RRDB_decoder = nn.Sequential(*[RRDB(nf, gc) for _ in range(nb)])
        self.conv_last = nn.Conv2d(nf, ou_ch, 3, 1, 1, bias=False)
        _kaiming_fanout_(self)

    @staticmethod
    def _rdbs_of(seq) -> List[ResidualDenseBlock_5]:
        out = []
        for rr in seq:
            out += [rr.RDB1, rr.RDB2, rr.RDB3]
        return out

    def forward(self, x):
        used = [p for name, p in self.named_parameters() if not name.startswith("trunk_conv.")]
        return _NetFn.apply(self, x, *used)

    def _forward_impl(self, x, st):
        dt, dev = act_dtype(), x.device
        n, c, h, w = x.shape
        nf, ctot = self.nf, self.nf + 4 * self.gc
        x_in = _io_buf(n, h, w, c, dt, dev)
        ops.nchw_to_nhwc(x, x_in)
        enc, dec = self._rdbs_of(self.RRDB_encoder), self._rdbs_of(self.RRDB_decoder)
        be = [ops.new_buf(n, h, w, ctot, dt, dev) for _ in enc]
        bd = [ops.new_buf(n, h, w, ctot, dt, dev) for _ in dec]
        tmp = Slice(ops.new_buf(n, h, w, nf, dt, dev))
        fea2 = Slice(ops.new_buf(n, h, w, nf, dt, dev))
        self._fprop(self.conv_first, x_in, Slice(be[0], 0, nf))
        self._chain_forward(enc, be, tmp)
        ops.add(Slice(be[0], 0, nf), tmp, Slice(bd[0], 0, nf))          # fea = fea + encoder(fea)
        self._chain_forward(dec, bd, tmp)
        ops.add(Slice(bd[0], 0, nf), tmp, fea2)                          # fea = fea + decoder(fea)
        out = _io_buf(n, h, w, self.conv_last.out_channels, dt, dev)
        self._fprop(self.conv_last, fea2, out)
        st["x_in"], st["be"], st["bd"], st["fea2"] = x_in, be, bd, fea2
        return ops.nhwc_to_nchw(out)

    def _backward_impl(self, st, grad_out, sink, want, need_dx):
        dt, dev = act_dtype(), grad_out.device
        W = lambda p: p is not None and want.get(id(p), False)
        x_in, be, bd, fea2 = st["x_in"], st["be"], st["bd"], st["fea2"]
        n, h, w, nf, ctot = x_in.n, x_in.h, x_in.w, self.nf, self.nf + 4 * self.gc
        enc, dec = self._rdbs_of(self.RRDB_encoder), self._rdbs_of(self.RRDB_decoder)
        d_o = _io_buf(n, h, w, self.conv_last.out_channels, dt, dev)
        ops.nchw_to_nhwc(grad_out, d_o)
        self._wgrad(self.conv_last, fea2, d_o, sink, W(self.conv_last.weight), False)
        Dbuf = [ops.new_buf(n, h, w, ctot, dt, dev) for _ in range(4)]
        g2 = Slice(ops.new_buf(n, h, w, nf, dt, dev))                   # d fea2 (kept: the chain recycles its slots)
        self._dgrad(self.conv_last, d_o, g2)
        Slice(Dbuf[(len(dec) - 1) % 4], 0, nf).view().copy_(g2.view())   # = d decoder-out
        through = Slice(ops.new_buf(n, h, w, nf, dt, dev))
        g1 = Slice(ops.new_buf(n, h, w, nf, dt, dev))
        self._chain_backward(dec, bd, Dbuf, through, sink, W)
        ops.add(through, g2, g1)                                          # d fea1 = skip + through the decoder
        Ebuf = [ops.new_buf(n, h, w, ctot, dt, dev) for _ in range(4)]
        ge = Slice(Ebuf[(len(enc) - 1) % 4], 0, nf)
        ge.view().copy_(g1.view())                                       # d encoder-out = d fea1
        self._chain_backward(enc, be, Ebuf, through, sink, W)
        d_fea = Slice(ops.new_buf(n, h, w, nf, dt, dev))
        ops.add(through, g1, d_fea)
        self._wgrad(self.conv_first, x_in, d_fea, sink, W(self.conv_first.weight), W(self.conv_first.bias))
        if not need_dx:
            return None
        dx = _io_buf(n, h, w, x_in.c, dt, dev)
        self._dgrad(self.conv_first, d_fea, dx)
        return ops.nhwc_to_nchw(dx)


# ------------------------------------------------------------------------------------------
# Plain conv stacks of the cascaded trainers (reference src/model/espcn.py, src/model/srcnn.py)
# ------------------------------------------------------------------------------------------

class _SeqConvNet(_NetBase):
    """conv (+ReLU) / PixelShuffle sequences.  ``_layers()`` -> [("conv", module, relu: bool) | ("shuffle", r)]."""

    def forward(self, x):
        return _NetFn.apply(self, x, *_flat_params(self))

    def _forward_impl(self, x, st):
        dt, dev = act_dtype(), x.device
        n, c, h, w = x.shape
        cur = _io_buf(n, h, w, c, dt, dev)
        ops.nchw_to_nhwc(x, cur)
        acts = [cur]
        for layer in self._layers():
            if layer[0] == "conv":
                conv, relu = layer[1], layer[2]
                nxt = _io_buf(n, cur.h, cur.w, conv.out_channels, dt, dev)
                self._fprop(conv, cur, nxt, act=(0.0 if relu else None))
            else:
                r = layer[1]
                nxt = _io_buf(n, cur.h * r, cur.w * r, cur.c // (r * r), dt, dev)
                ops.pixel_shuffle(cur, nxt, r)
            acts.append(nxt)
            cur = nxt
        st["acts"] = acts
        return ops.nhwc_to_nchw(cur)

    def _backward_impl(self, st, grad_out, sink, want, need_dx):
        dt, dev = act_dtype(), grad_out.device
        W = lambda p: p is not None and want.get(id(p), False)
        acts, layers = st["acts"], self._layers()
        n = acts[0].n
        g = _io_buf(n, acts[-1].h, acts[-1].w, acts[-1].c, dt, dev)
        ops.nchw_to_nhwc(grad_out, g)
        if layers[-1][0] == "conv" and layers[-1][2]:              # trailing ReLU (SRCNN): mask the incoming gradient
            g.view().mul_((acts[-1].view() > 0).to(g.dtype))
        for i in range(len(layers) - 1, -1, -1):
            layer, src = layers[i], acts[i]
            if i == 0 and not need_dx and layer[0] != "conv":
                return None
            prev_relu = i > 0 and layers[i - 1][0] == "conv" and layers[i - 1][2]
            if layer[0] == "conv":
                conv = layer[1]
                self._wgrad(conv, src, g, sink, W(conv.weight), W(conv.bias))
                if i == 0 and not need_dx:
                    return None
                nxt = _io_buf(n, src.h, src.w, src.c, dt, dev)
                if prev_relu:
                    self._dgrad(conv, g, nxt, mask=src, mask_slope=0.0)
                else:
                    self._dgrad(conv, g, nxt)
            else:
                nxt = _io_buf(n, src.h, src.w, src.c, dt, dev)
                ops.pixel_shuffle(g, nxt, layer[1], adjoint=True)
                if prev_relu:
                    nxt.view().mul_((src.view() > 0).to(nxt.dtype))
            g = nxt
        return ops.nhwc_to_nchw(g)


class ESPCN(_SeqConvNet):
    """ESPCN(in_ch, ou_ch, upscale_factor, base_kernel): 5x5, 3x3, 3x3 (ReLU) -> 3x3 -> PixelShuffle -> 3x3."""

    def __init__(self, in_ch=3, ou_ch=3, upscale_factor=2, base_kernel=64):
        super().__init__()
        k = [int(x * base_kernel) for x in [1, 1, 1 / 2]]
        self.up = upscale_factor
        self.conv1 = nn.Conv2d(in_ch, k[0], kernel_size=5, stride=1, padding=2)
        self.conv2 = nn.Conv2d(k[0], k[1], kernel_size=3, stride=1, padding=1)
        self.conv3 = nn.Conv2d(k[1], k[2], kernel_size=3, stride=1, padding=1)
        self.conv4 = nn.Conv2d(k[2], base_kernel * upscale_factor ** 2, kernel_size=3, stride=1, padding=1)
        self.conv5 = nn.Conv2d(base_kernel, ou_ch, kernel_size=3, stride=1, padding=1)
        _kaiming_fanout_(self)

    def _layers(self):
        return [("conv", self.conv1, True), ("conv", self.conv2, True), ("conv", self.conv3, True),
                ("conv", self.conv4, False), ("shuffle", self.up), ("conv", self.conv5, False)]


class SRCNN(_SeqConvNet):
    """SRCNN(in_ch, ou_ch, upscale_factor, base_kernel): 9x9, 1x1, 5x5, each followed by ReLU (no upsampling)."""

    def __init__(self, in_ch=3, ou_ch=3, upscale_factor=2, base_kernel=64):
        super().__init__()
        k = [int(x * base_kernel) for x in [1, 1 / 2]]
        self.up = upscale_factor
        self.conv1 = nn.Conv2d(in_ch, k[0], kernel_size=9, stride=1, padding=4)
        self.conv2 = nn.Conv2d(k[0], k[1], kernel_size=1, stride=1, padding=0)
        self.conv3 = nn.Conv2d(k[1], ou_ch, kernel_size=5, stride=1, padding=2)

    def _layers(self):
        return [("conv", self.conv1, True), ("conv", self.conv2, True), ("conv", self.conv3, True)]


# ------------------------------------------------------------------------------------------
# NLayerDiscriminator (reference model.py:595-639), BatchNorm2d norm layer
# ------------------------------------------------------------------------------------------

class NLayerDiscriminator(_NetBase):
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d):
        super().__init__()
        if norm_layer is not nn.BatchNorm2d:
            raise NotImplementedError("srcgan_b200.NLayerDiscriminator: only nn.BatchNorm2d is implemented "
                                      "(the only norm layer the reference's trainers use, train.py:169-180)")
        seq: List[nn.Module] = [nn.Conv2d(input_nc, ndf, 4, 2, 1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for i in range(1, n_layers):
            prev, mult = mult, min(2 ** i, 8)
            seq += [nn.Conv2d(ndf * prev, ndf * mult, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * mult),
                    nn.LeakyReLU(0.2, True)]
        prev, mult = mult, min(2 ** n_layers, 8)
        seq += [nn.Conv2d(ndf * prev, ndf * mult, 4, 1, 1, bias=False), nn.BatchNorm2d(ndf * mult),
                nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(ndf * mult, 1, 4, 1, 1)]
        self.model = nn.Sequential(*seq)     # same indices => same state_dict keys as the reference

    def _plan(self):
        """[(conv, bn or None, has_act)] in execution order."""
        mods = list(self.model)
        plan, i = [], 0
        while i < len(mods):
            conv = mods[i]; i += 1
            bn = None
            if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
                bn = mods[i]; i += 1
            act = False
            if i < len(mods) and isinstance(mods[i], nn.LeakyReLU):
                act = True; i += 1
            plan.append((conv, bn, act))
        return plan

    def forward(self, input):
        return _NetFn.apply(self, input, *_flat_params(self))

    def _forward_impl(self, x, st):
        dt, dev = act_dtype(), x.device
        n, c, h, w = x.shape
        cur = _io_buf(n, h, w, c, dt, dev)
        ops.nchw_to_nhwc(x, cur)
        ins, ys, outs, stats = [], [], [], []
        training = self.training
        for conv, bn, act in self._plan():
            k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
            ho, wo = (cur.h + 2 * p - k) // s + 1, (cur.w + 2 * p - k) // s + 1
            if ho < 1 or wo < 1:
                raise RuntimeError("NLayerDiscriminator: input %dx%d too small" % (h, w))
            ins.append(cur)
            y = _io_buf(n, ho, wo, conv.out_channels, dt, dev)
            if bn is None:
                self._fprop(conv, cur, y, act=(LRELU if act else None))
                ys.append(None); stats.append(None)
                cur = y
            else:
                self._fprop(conv, cur, y)
                z = Slice(ops.new_buf(n, ho, wo, conv.out_channels, dt, dev))
                use_batch = training or bn.running_mean is None
                sm, si = ops.bn_forward(y, z, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch,
                                        LRELU if act else 1.0, bn.momentum if bn.momentum is not None else 0.1, bn.eps)
                if training and bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
                ys.append(y); stats.append((sm, si, use_batch))
                cur = z
            outs.append(cur)
        st["ins"], st["ys"], st["outs"], st["stats"] = ins, ys, outs, stats
        return ops.nhwc_to_nchw(cur)

    def _backward_impl(self, st, grad_out, sink, want, need_dx):
        dt, dev = act_dtype(), grad_out.device
        ins, ys, outs, stats = st["ins"], st["ys"], st["outs"], st["stats"]
        W = lambda p: p is not None and want.get(id(p), False)
        plan = self._plan()
        n = ins[0].n
        g = _io_buf(n, outs[-1].h, outs[-1].w, outs[-1].c, dt, dev)
        ops.nchw_to_nhwc(grad_out, g)
        # g is always the gradient w.r.t. the *conv output* of layer i when we reach its wgrad
        for i in range(len(plan) - 1, -1, -1):
            conv, bn, act = plan[i]
            if bn is not None:
                sm, si, use_batch = stats[i]
                dg, acc_g = sink.slot(bn.weight, W(bn.weight))
                db, acc_b = sink.slot(bn.bias, W(bn.bias))
                ops.bn_backward(g, outs[i], ys[i], g, bn.weight, sm, si, LRELU if act else 1.0, use_batch, dg, db,
                                acc_g or acc_b, beta=bn.bias)
            self._wgrad(conv, ins[i], g, sink, W(conv.weight), W(conv.bias))
            if i == 0 and not need_dx:
                return None
            src = ins[i]
            nxt = _io_buf(n, src.h, src.w, src.c, dt, dev)
            # the previous layer's LeakyReLU (no BN) is folded into this dgrad's mask epilogue
            prev_plain_act = i > 0 and plan[i - 1][1] is None and plan[i - 1][2]
            if prev_plain_act:
                self._dgrad(conv, g, nxt, mask=outs[i - 1], mask_slope=LRELU)
            else:
                self._dgrad(conv, g, nxt)
            g = nxt
        return ops.nhwc_to_nchw(g)
