"""Host-side orchestration of the sm_100a kernels (through the C ABI in _lib.py).

Nothing here computes on the CPU and there is no torch-op fallback for a kernel: torch is
used for device memory (workspaces, packed weights), stream handles and tiny D2D copies.

Data layout in HBM: every activation is NHWC fp32, ``[B, h, w, C]`` with an
explicit leading dimension so that the reference's ``torch.cat([...], dim=1)`` becomes
"several producers write channel slices of one buffer" or "one conv reads several slices".
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L

F32 = torch.float32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class View:
    """A (batch-range, channel-range) slice of an NHWC fp32 tensor ``t`` [B, h, w, ld]."""
    __slots__ = ("t", "b0", "b", "h", "w", "c0", "c", "ld", "ptr")

    def __init__(self, t: torch.Tensor, b0: int = 0, b: Optional[int] = None, c0: int = 0, c: Optional[int] = None):
        assert t.dtype == F32 and t.is_contiguous() and t.dim() == 4
        self.t = t
        self.h, self.w, self.ld = t.shape[1], t.shape[2], t.shape[3]
        self.b0, self.b = b0, (t.shape[0] - b0 if b is None else b)
        self.c0, self.c = c0, (self.ld - c0 if c is None else c)
        self.ptr = t.data_ptr() + 4 * (b0 * self.h * self.w * self.ld + c0)

    def ch(self, c0: int, c1: int) -> "View":
        assert 0 <= c0 < c1 <= self.c
        return View(self.t, self.b0, self.b, self.c0 + c0, c1 - c0)

    def rows(self, b0: int, b1: int) -> "View":
        assert 0 <= b0 < b1 <= self.b
        return View(self.t, self.b0 + b0, b1 - b0, self.c0, self.c)

    @property
    def full_rows(self) -> bool:
        return self.b0 == 0 and self.b == self.t.shape[0]

    @property
    def npix(self):
        return self.b * self.h * self.w


class PlanesOnly:
    """Shape carrier for a conv source that exists only as operand planes (no fp32 tensor)."""
    __slots__ = ("b", "h", "w", "c", "ld", "ptr")

    def __init__(self, b, h, w, c, ld):
        self.b, self.h, self.w, self.c, self.ld, self.ptr = b, h, w, c, ld, 0


class PackedConv:
    """Weights of one conv in the kernel layout [kh*kw][cin][cout_pad] + epilogue vectors."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]], stride=1,
                 pad=(0, 0), bn=None, out_scale: Optional[torch.Tensor] = None, mult: float = 1.0):
        w = torch.cat([x.detach().to(F32) for x in weights], 0)           # concat along cout
        dev = w.device
        self.w_oihw = w
        self._tc = None
        self.cout, self.cin, self.kh, self.kw = w.shape
        self.stride, self.pad_h, self.pad_w = stride, pad[0], pad[1]
        self.cout_pad = (self.cout + 3) // 4 * 4
        packed = torch.zeros(self.kh * self.kw, self.cin, self.cout_pad, device=dev, dtype=F32)
        packed[:, :, : self.cout] = w.permute(2, 3, 1, 0).reshape(self.kh * self.kw, self.cin, self.cout)
        self.w = packed.contiguous()
        bias = torch.cat([(torch.zeros(x.shape[0], device=dev) if b is None else b.detach().to(F32))
                          for x, b in zip(weights, biases)])
        scale = None
        shift = bias
        if bn is not None:                      # eval BatchNorm folded into the epilogue affine
            g, beta, rm, rv, eps = bn
            inv = (g / torch.sqrt(rv + eps)).to(F32)
            scale = inv
            shift = (bias - rm) * inv + beta
        if out_scale is not None:               # ZeroConv2d: (conv + b) * exp(3*scale)
            scale = out_scale if scale is None else scale * out_scale
            shift = shift * out_scale
        if mult != 1.0:                         # mask head: 0.25 * (conv + b)
            scale = torch.full_like(shift, mult) if scale is None else scale * mult
            shift = shift * mult
        self.scale = None if scale is None else scale.contiguous()
        self.shift = shift.contiguous()
        self.has_bias = any(b is not None for b in biases) or bn is not None

    def as_1x1(self) -> "PackedConv":
        """The same filter as a 1x1 conv over im2col'd input with K ordered (ky, kx, cin)."""
        if getattr(self, "_as1x1", None) is None:
            w = self.w_oihw.permute(0, 2, 3, 1).reshape(self.cout, -1, 1, 1).contiguous()
            pc = PackedConv.__new__(PackedConv)
            pc.w_oihw, pc._tc, pc._as1x1 = w, None, None
            pc.cout, pc.cin, pc.kh, pc.kw = w.shape
            pc.stride, pc.pad_h, pc.pad_w, pc.cout_pad = 1, 0, 0, self.cout_pad
            pc.w = None
            pc.scale, pc.shift, pc.has_bias = self.scale, self.shift, self.has_bias
            self._as1x1 = pc
        return self._as1x1

    def as_taps1x1(self) -> "PackedConv":
        """A 3x3 filter with few outputs as a 1x1 conv with 9*cout outputs (row = tap*cout + o); bias, scale and
        activation are applied after the nine taps are summed (accflow_tapsum3x3_f32)."""
        if getattr(self, "_astaps", None) is None:
            w = self.w_oihw.permute(2, 3, 0, 1).reshape(self.kh * self.kw * self.cout, self.cin, 1, 1)
            if w.shape[0] % 4:       # zero rows up to a multiple of 4: every output group takes the epilogue's vector path
                w = torch.cat([w, w.new_zeros(4 - w.shape[0] % 4, self.cin, 1, 1)])
            w = w.contiguous()
            pc = PackedConv.__new__(PackedConv)
            pc.w_oihw, pc._tc, pc._as1x1, pc._astaps = w, None, None, None
            pc.cout, pc.cin, pc.kh, pc.kw = w.shape
            pc.stride, pc.pad_h, pc.pad_w, pc.cout_pad = 1, 0, 0, (pc.cout + 3) // 4 * 4
            pc.w, pc.scale, pc.shift, pc.has_bias = None, None, None, False
            self._astaps = pc
        return self._astaps

    def tc_weights(self, fp16x2=False) -> "L.TcWeights":
        """16-bit operand planes [planes][taps][cout][k_pitch] for accflow_conv2d_tc; ``fp16x2``: False = bf16 planes,
        True = fp16 hi + scaled lo, "fp16" = one fp16 plane."""
        if self._tc is None:
            self._tc = {}
        if fp16x2 not in self._tc:
            taps = self.kh * self.kw
            pitch = Kernels.plane_pitch(self.cin)          # 128-byte aligned rows, like the activation planes
            wt = torch.zeros(taps, self.cout, pitch, device=self.w_oihw.device, dtype=F32)
            wt[:, :, : self.cin] = self.w_oihw.permute(2, 3, 0, 1).reshape(taps, self.cout, self.cin)
            planes = split_planes_torch(wt, fp16x2)
            tw = L.TcWeights(planes.data_ptr(), planes.shape[0], self.cout, self.cin, pitch, taps)
            self._tc[fp16x2] = (tw, planes)
        return self._tc[fp16x2][0]


def split_planes_torch(x: torch.Tensor, fp16x2: bool = False) -> torch.Tensor:
    """x (fp32) -> stacked 16-bit planes (one-off weight packing): bf16 p0, p1, p2 with x ~= p0 + p1 + p2,
    or fp16 (hi, lo * 2^11) when ``fp16x2``."""
    if fp16x2 == "fp16":
        return x.clamp(-65504.0, 65504.0).to(torch.float16)[None].contiguous()
    if fp16x2:
        x = x.clamp(-65504.0, 65504.0)           # same saturation as the device-side split (common.cuh sat_fp16)
        hi = x.to(torch.float16)
        lo = ((x - hi.float()) * 2048.0).to(torch.float16)
        return torch.stack([hi, lo]).contiguous()
    p0 = x.to(torch.bfloat16)
    r1 = x - p0.float()
    p1 = r1.to(torch.bfloat16)
    p2 = (r1 - p1.float()).to(torch.bfloat16)
    return torch.stack([p0, p1, p2]).contiguous()


class Kernels:
    """Thin typed wrappers over the C ABI.  One instance per device."""

    MODES = ("fp32", "bf16x3", "fp16x2", "bf16", "fp16")
    NPROD = {"bf16x3": 6, "fp16x2": 3, "bf16": 1, "fp16": 2}           # accflow_conv2d_tc's nprod codes
    NPLANES = {"bf16x3": 3, "fp16x2": 2, "bf16": 1, "fp16": 1, "fp32": 0}
    PLANE_FMT = {"bf16x3": 3, "fp16x2": 2, "bf16": 1, "fp16": 4, "fp32": 0}     # format codes of include/accflow_b200.h

    def __init__(self, device: torch.device, precision: str = "fp32"):
        assert precision in self.MODES, precision
        self.device = device
        self.precision = precision       # fp32: FFMA kernels; bf16x3 / bf16: tcgen05 kernels (6 / 1 products)
        L.load()
        self._ws: Dict[tuple, torch.Tensor] = {}
        # bf16 planes that accompany fp32 activations consumed by the tensor-core convs:
        # root data_ptr -> planes tensor [3, B, h, w, Cp] and the channel ranges whose planes are stale
        self._planes: Dict[int, torch.Tensor] = {}
        self._stale: Dict[int, list] = {}
        self._roots: Dict[int, torch.Tensor] = {}   # keeps the fp32 root alive so its address cannot be recycled
        self.profile = None      # bench.py: list of (start_event, end_event, flops) per conv launch

    # ---- workspace -----------------------------------------------------------------------
    def buf(self, name: str, *shape, zero: bool = False) -> torch.Tensor:
        key = (name,) + tuple(shape)
        t = self._ws.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(*shape, device=self.device, dtype=F32)
            self._ws[key] = t
        return t

    def view(self, name: str, b: int, h: int, w: int, c: int) -> View:
        return View(self.buf(name, b, h, w, c))

    # ---- bf16 planes bookkeeping ---------------------------------------------------------
    @property
    def tc(self) -> bool:
        return self.precision != "fp32"

    @property
    def nplanes(self) -> int:
        """Number of 16-bit planes per operand (allocation, strides)."""
        return self.NPLANES[self.precision]

    @property
    def plane_fmt(self) -> int:
        """Plane format code handed to every kernel that writes operand planes."""
        return self.PLANE_FMT[self.precision]

    def wrote(self, v: Optional[View]):
        """A non-tensor-core kernel wrote ``v``: its planes (if any exist) are stale."""
        if v is None or not self.tc:
            return
        key = v.t.data_ptr()
        if key in self._planes:
            rng = (v.c0, v.c0 + v.c)
            st = self._stale[key]
            if rng not in st:
                st.append(rng)

    # Plane pitch: a pixel's channels start on a 128-byte line (64 elements), so every 64-channel K block a TMA box
    # fetches is one whole L2 line per pixel instead of two partial ones (324 -> 384, 104 -> 128, 96 -> 128, 48 -> 64).
    PLANE_ALIGN = int(os.environ.get("ACCFLOW_PLANE_ALIGN", "64"))

    @classmethod
    def plane_pitch(cls, ld: int) -> int:
        a = cls.PLANE_ALIGN if ld >= 32 else 8
        return (ld + a - 1) // a * a

    def _plane_geom(self, v: View):
        cp = self.plane_pitch(v.ld)
        return cp, v.t.shape[0] * v.h * v.w * cp

    def planes_ptr(self, v: View, create: bool):
        """-> (ptr of the slice's first element in plane 0, pitch, plane_stride) or None."""
        key = v.t.data_ptr()
        pl = self._planes.get(key)
        if pl is None:
            if not create:
                return None
            cp, _ = self._plane_geom(v)
            pl = torch.empty(self.nplanes, v.t.shape[0], v.h, v.w, cp, device=self.device, dtype=torch.bfloat16)
            self._planes[key] = pl
            self._stale[key] = [(0, v.ld)]
            self._roots[key] = v.t
        cp, stride = self._plane_geom(v)
        return pl.data_ptr() + 2 * (v.b0 * v.h * v.w * cp + v.c0), cp, stride

    def release_planes(self):
        """Drop all plane buffers (they are rebuilt on demand)."""
        self._planes.clear()
        self._stale.clear()
        self._roots.clear()

    def ensure_planes(self, v: View):
        """Make the planes of ``v`` current (split pass over the stale channel ranges it overlaps)."""
        ptr, cp, stride = self.planes_ptr(v, create=True)
        key = v.t.data_ptr()
        st = self._stale[key]
        lo, hi = v.c0, v.c0 + v.c
        for rng in [r for r in st if r[0] < hi and r[1] > lo]:
            r0, r1 = rng
            rows = v.t.shape[0] * v.h * v.w
            L.call("accflow_split_bf16_planes", v.t.data_ptr() + 4 * r0, rows, r1 - r0, v.ld, r1 - r0, cp, stride,
                   self.plane_fmt, self._planes[key].data_ptr() + 2 * r0, _stream())
            st.remove(rng)
        return ptr, cp, stride

    def _fresh(self, v: View):
        """The tensor-core epilogue just wrote the planes of ``v``."""
        if not v.full_rows:
            return
        key = v.t.data_ptr()
        lo, hi = v.c0, v.c0 + v.c
        self._stale[key] = [r for r in self._stale[key] if not (r[0] >= lo and r[1] <= hi)]

    # ---- convolution ---------------------------------------------------------------------
    def conv(self, pc: PackedConv, srcs: Sequence[View], out: Optional[View] = None, act=L.ACT_NONE, alpha=1.0,
             act_split=0, act2=L.ACT_NONE, out2: Optional[View] = None, residual: Optional[View] = None,
             post_relu=False, epilogue=L.EPI_STORE, h: Optional[View] = None, z: Optional[View] = None,
             weight_ptr: Optional[int] = None, weight_batch_stride=0, cout=None, cout_pad=None,
             use_affine=True, tc_b: Optional["L.TcWeights"] = None, tc_src_planes=None, planes_only=False,
             emit_planes=True, pool_w=0, pre_add: Optional[View] = None, row_stats: Optional[torch.Tensor] = None,
             tc_out_planes=None, out_h: int = 0, pre_mod: int = 0, tile_order: int = 0):
        """``planes_only``: the output (``out``; ``out2`` for the GRU z|r epilogue) is read by tensor-core
        convolutions only, so in the tensor-core modes its fp32 copy is not written (half the store bytes).
        ``emit_planes=False``: the next reader is not a tensor-core conv (InstanceNorm), so no planes are written.
        ``pre_add``: fp32 slice added before the activation / gate math (the hoisted GRU ``inp`` term); with
        ``pre_mod`` > 0 it holds ``pre_mod`` samples and sample s reads sample s % pre_mod (pairs that share their
        first frame share the term).  ``tile_order``: L.TILES_FORWARD / L.TILES_REVERSE - walk the output tiles
        opposite to the producer of the input (it then starts on what is still in L2); 0 = alternate per launch."""
        d = L.ConvDesc()
        d.tile_order = tile_order
        cin = 0
        for k, s in enumerate(srcs):
            d.src[k], d.src_c[k], d.src_ld[k] = s.ptr, s.c, s.ld
            cin += s.c
        assert cin == pc.cin, f"conv expects {pc.cin} input channels, got {cin}"
        s0 = srcs[0]
        d.nsrc, d.batch, d.in_h, d.in_w = len(srcs), s0.b, s0.h, s0.w
        tc = self.tc
        if not tc:
            d.weight = pc.w.data_ptr() if weight_ptr is None else weight_ptr
        d.weight_batch_stride = weight_batch_stride
        d.kh, d.kw, d.stride, d.pad_h, d.pad_w = pc.kh, pc.kw, pc.stride, pc.pad_h, pc.pad_w
        d.cout = pc.cout if cout is None else cout
        d.cout_pad = pc.cout_pad if cout_pad is None else cout_pad
        d.alpha = alpha
        if use_affine:
            d.scale = None if pc.scale is None else pc.scale.data_ptr()
            d.shift = pc.shift.data_ptr() if pc.has_bias else None
        d.act, d.act_split, d.act2 = act, act_split, act2
        if residual is not None:
            d.residual, d.res_ld = residual.ptr, residual.ld
        d.post_relu, d.epilogue, d.pool_w = int(post_relu), epilogue, pool_w
        if out is not None:
            d.out, d.out_ld = out.ptr, out.ld
        if out2 is not None:
            d.out2, d.out2_ld = out2.ptr, out2.ld
        if h is not None:
            d.h, d.h_ld = h.ptr, h.ld
        if z is not None:
            d.z, d.z_ld = z.ptr, z.ld
        if pre_add is not None:
            assert pre_add.c == d.cout and pre_add.b == (pre_mod or s0.b) and s0.b % (pre_mod or s0.b) == 0
            d.pre_add, d.pre_ld, d.pre_mod = pre_add.ptr, pre_add.ld, pre_mod
        if row_stats is not None:
            d.row_stats = row_stats.data_ptr()
        d.out_h = out_h
        if tc:
            io = L.TcIO()
            for k, sv in enumerate(srcs):
                io.src_planes[k], io.src_pitch[k], io.src_plane_stride[k] = (
                    tc_src_planes[k] if tc_src_planes is not None else self.ensure_planes(sv))
            written = []
            if epilogue in (L.EPI_STORE_POOL, L.EPI_ROWSTATS, L.EPI_STORE_T):
                targets = ()
            elif epilogue == L.EPI_STORE:
                targets = (("out", out), ("out2", out2 if act_split else None))
            elif epilogue == L.EPI_GRU_ZR:
                targets = (("out2", out2),)
            else:
                targets = (("h", h),)
            for name, tv in targets:
                if tv is None:
                    continue
                got = self.planes_ptr(tv, create=False) if emit_planes else None
                if got is not None and tv.c0 % 4 == 0:
                    setattr(io, name + "_planes", got[0])
                    setattr(io, name + "_pitch", got[1])
                    setattr(io, name + "_plane_stride", got[2])
                    written.append(tv)
                else:
                    self.wrote(tv)
            if planes_only and written and written[0].full_rows:      # the planes carry the result; drop the fp32 store
                if epilogue == L.EPI_STORE and not act_split and written[0] is out:
                    d.out = None
                elif epilogue == L.EPI_GRU_ZR and written[0] is out2:
                    d.out2 = None
            if tc_out_planes is not None:          # caller-owned destination planes (no fp32 twin tensor)
                io.out_planes, io.out_pitch, io.out_plane_stride = tc_out_planes
            tw = tc_b if tc_b is not None else pc.tc_weights(
                {"fp16x2": True, "fp16": "fp16"}.get(self.precision, False))
            args = ("accflow_conv2d_tc", C.byref(d), C.byref(io), C.byref(tw), self.NPROD[self.precision], _stream())
        else:
            args = ("accflow_conv2d_f32", C.byref(d), _stream())
            written = []
        if self.profile is None:
            L.call(*args)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            oh = out_h or (s0.h + 2 * pc.pad_h - pc.kh) // pc.stride + 1
            ow = (s0.w + 2 * pc.pad_w - pc.kw) // pc.stride + 1
            e0.record()
            L.call(*args)
            e1.record()
            macs = getattr(pc, "algo_macs_per_pixel", None) or d.cout * cin * pc.kh * pc.kw     # algorithmic, not padded
            self.profile.append((e0, e1, 2.0 * s0.b * oh * ow * macs))
        for tv in written:
            self._fresh(tv)

    def _out_planes(self, v: View):
        """Planes pointer triple for a kernel that can emit planes itself, or (None, 0, 0)."""
        got = self.planes_ptr(v, create=False) if self.tc else None
        return got if got is not None else (None, 0, 0)

    def _done(self, v: View, emitted: bool):
        if emitted:
            self._fresh(v)
        else:
            self.wrote(v)

    def conv_smallc(self, x_ptr: int, nchw: bool, batch, cin, h, w, pc: PackedConv, act, out: View):
        pl = self._out_planes(out)
        L.call("accflow_conv_smallc_f32", x_ptr, int(nchw), batch, cin, h, w, pc.w.data_ptr(),
               None if pc.scale is None else pc.scale.data_ptr(), pc.shift.data_ptr(), pc.kh, pc.stride, pc.cout,
               act, out.ptr, out.ld, pl[0], pl[1], pl[2], self.plane_fmt, _stream())
        self._done(out, pl[0] is not None)

    def flow_conv7(self, tag: str, flow: torch.Tensor, batch: int, h: int, w: int, pc: PackedConv, out: View,
                   planes_only=False, tile_order: int = 0):
        """relu(conv7x7(flow)) for a 2-channel flow field [batch, h*w, 2] (raft/update.py:92, AccFlow_.py:62)."""
        if not self.tc:
            self.conv_smallc(flow.data_ptr(), False, batch, 2, h, w, pc, L.ACT_RELU, out)
            return
        patch = self.view(tag + ".fpatch", batch, h, w, 104)
        pl = self.planes_ptr(patch, create=True)
        L.call("accflow_flow_patch_f32", flow.data_ptr(), batch, h, w, None, patch.ld, pl[0], pl[1], pl[2],
               self.plane_fmt, _stream())          # planes only: nothing reads the fp32 patch
        self._stale[patch.t.data_ptr()] = []
        self.conv(pc.as_1x1(), [patch.ch(0, 98)], out, act=L.ACT_RELU, planes_only=planes_only, tile_order=tile_order)

    def conv_smallcout(self, pc: PackedConv, x: View, out: View, act=L.ACT_NONE, accum: Optional[torch.Tensor] = None,
                       accum_ld: int = 0, tile_order: int = 0):
        """3x3 conv with <= 4 output channels; ``accum`` (optional, [pixels, accum_ld]) += result.
        Tensor-core modes: a 1x1 conv with 9*cout outputs (each activation read once) + a 9-tap sum;
        fp32 mode: the FFMA bandwidth kernel."""
        assert pc.kh == 3 and pc.kw == 3 and pc.stride == 1 and pc.cout <= 4 and pc.cout_pad == 4 and x.c == pc.cin
        scale = None if pc.scale is None else pc.scale.data_ptr()
        if self.tc:
            n9 = 9 * pc.cout
            t = self.view(f"tapsum{n9}", x.b, x.h, x.w, (n9 + 3) // 4 * 4)
            self.conv(pc.as_taps1x1(), [x], t, tile_order=tile_order)   # (rows 9*cout.. of the padded filter are zero)
            L.call("accflow_tapsum3x3_f32", t.ptr, t.ld, x.b, x.h, x.w, pc.cout, scale, pc.shift.data_ptr(), act,
                   out.ptr, out.ld, None if accum is None else accum.data_ptr(), accum_ld, _stream())
        else:
            L.call("accflow_conv3x3_smallcout_f32", x.ptr, x.ld, x.b, x.h, x.w, x.c, pc.w.data_ptr(), scale,
                   pc.shift.data_ptr(), pc.cout, act, out.ptr, out.ld, _stream())
            if accum is not None:
                assert accum_ld == out.ld == pc.cout
                L.call("accflow_axpy_f32", accum.data_ptr(), out.ptr, 1.0, x.b * x.h * x.w * pc.cout, _stream())
        self.wrote(out)

    def corr_lookup(self, lv, radius: int, coords: torch.Tensor, out: View, flow: torch.Tensor, mf_tail: View,
                    planes_only=False):
        """CorrBlock.__call__ for all four levels; also emits flow = coords - grid."""
        pl, tl = self._out_planes(out), self._out_planes(mf_tail)
        out_f32 = None if (planes_only and pl[0] is not None) else out.ptr      # convc1 reads the planes only
        L.call("accflow_corr_lookup_f32", lv[0].data_ptr(), lv[1].data_ptr(), lv[2].data_ptr(), lv[3].data_ptr(),
               out.b, out.h, out.w, radius, coords.data_ptr(), out_f32, out.ld, flow.data_ptr(), mf_tail.ptr, mf_tail.ld,
               pl[0], pl[1], pl[2], tl[0], tl[1], tl[2], self.plane_fmt, _stream())
        self._done(out, pl[0] is not None)
        self._done(mf_tail, tl[0] is not None)

    def instnorm(self, x: View, relu: bool, residual: Optional[View], post_relu: bool, out: View, eps=1e-5,
                 planes_only=False, emit_planes=True):
        """InstanceNorm2d (+ReLU, +residual, +ReLU).  Tensor-core modes: the apply pass also writes the operand
        planes of ``out`` (no split pass before the next conv); ``planes_only`` additionally drops the fp32
        copy when tensor-core convolutions are the only readers."""
        assert x.c == x.ld and out.c == out.ld
        hw = x.h * x.w
        chunks = L.call("accflow_instnorm_chunks", hw)
        partial = self.buf("in_partial", x.b * chunks * x.c * 2)
        stats = self.buf("in_stats", x.b * x.c * 2)
        pl = self.planes_ptr(out, create=True) if self.tc and emit_planes and out.full_rows and out.c % 8 == 0 else None
        if pl is None:
            L.call("accflow_instnorm_f32", x.ptr, x.b, hw, x.c, eps, int(relu),
                   None if residual is None else residual.ptr, int(post_relu), out.ptr, partial.data_ptr(),
                   stats.data_ptr(), _stream())
            self.wrote(out)
            return
        L.call("accflow_instnorm_planes_f32", x.ptr, x.b, hw, x.c, eps, int(relu),
               None if residual is None else residual.ptr, int(post_relu), None if planes_only else out.ptr,
               partial.data_ptr(), stats.data_ptr(), pl[0], pl[1], pl[2], self.plane_fmt, _stream())
        self._fresh(out)

    def gemm_nt(self, tag: str, a: View, b: View, out: View, alpha=1.0, pool_out: Optional[torch.Tensor] = None,
                pool_w: int = 0):
        """out[s, m, n] = alpha * sum_k a[s, m, k] * b[s, n, k]  (per sample s; b given row-major [n][k]).
        corr volume (raft/corr.py:47-55) and q k^T (gma/modules.py:66-73)."""
        B, K, N = a.b, a.c, b.h * b.w
        assert b.c == K and b.b == B
        if self.precision == "fp32":
            Np = _p4(N)
            bt = self.buf(tag + ".bt", B, K, Np, zero=True)
            self.transpose(b, bt, Np)
            self.conv(_Gemm(K, N, Np), [a], out, alpha=alpha, weight_ptr=bt.data_ptr(), weight_batch_stride=K * Np,
                      use_affine=False)
            return
        npl = self.nplanes
        pitch = (K + 7) // 8 * 8
        planes = self.buf16(tag + ".bpl", npl, B, N, pitch)
        L.call("accflow_split_bf16_planes", b.ptr, B * N, K, b.ld, K, pitch, B * N * pitch, self.plane_fmt, planes.data_ptr(),
               _stream())
        tw = L.TcWeights(planes.data_ptr(), npl, N, K, pitch, B)
        if pool_out is not None:      # fused first pyramid level (ACCFLOW_EPI_STORE_POOL)
            self.conv(_Gemm(K, N, N), [a], out, alpha=alpha, weight_batch_stride=1, use_affine=False, tc_b=tw,
                      epilogue=L.EPI_STORE_POOL, out2=View(pool_out.view(B, a.h, a.w, -1)), pool_w=pool_w)
            return
        self.conv(_Gemm(K, N, N), [a], out, alpha=alpha, weight_batch_stride=1, use_affine=False, tc_b=tw)

    def gemm_nn(self, tag: str, a: View, b: View, out: View, alpha=1.0, residual: Optional[View] = None):
        """out[s, m, n] = alpha * sum_k a[s, m, k] * b[s, k, n] (+ residual); b given as [k][n] (NHWC rows = k).
        attn @ v (gma/modules.py:108)."""
        B, K, N = a.b, a.c, b.c
        assert b.h * b.w == K and b.b == B
        if self.precision == "fp32":
            assert b.ld == N
            self.conv(_Gemm(K, N, N), [a], out, alpha=alpha, weight_ptr=b.ptr, weight_batch_stride=K * N,
                      residual=residual, use_affine=False)
            return
        npl = self.nplanes
        Kp = (K + 7) // 8 * 8
        bt = self.buf(tag + ".bt", B, N, Kp, zero=True)
        self.transpose(b, bt, Kp)
        planes = self.buf16(tag + ".bpl", npl, B, N, Kp)
        L.call("accflow_split_bf16_planes", bt.data_ptr(), B * N, K, Kp, K, Kp, B * N * Kp, self.plane_fmt, planes.data_ptr(),
               _stream())
        tw = L.TcWeights(planes.data_ptr(), npl, N, K, Kp, B)
        self.conv(_Gemm(K, N, N), [a], out, alpha=alpha, weight_batch_stride=1, residual=residual, use_affine=False,
                  tc_b=tw)

    def buf16(self, name: str, *shape) -> torch.Tensor:
        key = (name, "bf16") + tuple(shape)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(*shape, device=self.device, dtype=torch.bfloat16)
            self._ws[key] = t
        return t

    def transpose(self, x: View, out: torch.Tensor, out_ld: int):
        L.call("accflow_nhwc_transpose_f32", x.ptr, x.b, x.h * x.w, x.c, x.ld, out.data_ptr(), out_ld, _stream())

    def softmax_rows(self, t: torch.Tensor, rows: int, n: int):
        L.call("accflow_softmax_rows_f32", t.data_ptr(), rows, n, _stream())
        key = t.data_ptr()
        if key in self._planes:
            self._stale[key] = [(0, t.shape[-1])]


def _p4(n: int) -> int:
    return (n + 3) // 4 * 4


class GraphCache:
    """CUDA-graph replay of a whole forward pass.

    The eager pass issues ~1.7k kernel launches per clip batch through ctypes; replaying them as
    one captured graph removes the host from the critical path (SURVEY.md §7 hard part 5).
    Inputs are copied into static buffers, the captured outputs are cloned for the caller.
    Capture happens on the third call with a given key (two eager warm-ups first, so that every
    workspace / plane buffer exists and the plane-freshness state is in its steady cycle).
    """

    def __init__(self):
        self.entries = {}

    def run(self, key, inputs: Sequence[Optional[torch.Tensor]], fn):
        ent = self.entries.get(key)
        if ent is None:
            static_in = [None if t is None else torch.empty_like(t) for t in inputs]
            ent = self.entries[key] = {"in": static_in, "graph": None, "out": None, "warm": 0}
        for dst, src in zip(ent["in"], inputs):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        if ent["graph"] is None:
            if ent["warm"] < 2:
                ent["warm"] += 1
                return fn(*ent["in"])
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = L.call("accflow_launch_count", 0)
            # thread_local: under nn.DataParallel the other replicas' threads keep launching while this one captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = fn(*ent["in"])
            ent["graph"], ent["out"] = graph, out
            ent["launches"] = L.call("accflow_launch_count", 0) - n0
            L.call("accflow_launch_count_add", -ent["launches"])     # capture itself executed nothing
        ent["graph"].replay()
        L.call("accflow_launch_count_add", ent["launches"])
        out = ent["out"]
        if isinstance(out, (list, tuple)):
            return [o.clone() for o in out]
        return out.clone()


# ================================================================================ encoders
class EncoderPlan:
    """BasicEncoder (raft/extractor.py:137-225) with norm in {'instance','batch','none'}."""

    def __init__(self, sd, pfx: str, norm: str, out_dim: int):
        self.norm, self.out_dim, self.pfx = norm, out_dim, pfx

        def bn(name):
            if norm != "batch":
                return None
            return (sd[name + ".weight"], sd[name + ".bias"], sd[name + ".running_mean"], sd[name + ".running_var"], 1e-5)

        def pc(name, stride, pad, bnname=None):
            return PackedConv([sd[name + ".weight"]], [sd[name + ".bias"]], stride, (pad, pad),
                              bn(bnname) if bnname else None)

        w = sd[pfx + "conv1.weight"]                       # (64,3,7,7) -> [(ky*7+kx)*3 + c][64]
        self.stem = pc(pfx + "conv1", 2, 3, pfx + "norm1")
        self.stem.w = w.detach().to(F32).permute(2, 3, 1, 0).reshape(147, 64).contiguous()
        # tensor-core form (accflow_stem_rows_planes): 4 vertical taps over 48 channels k = tx*12 + c*4 + dy*2 + dx,
        # filter value w[o][c][2*ty + dy - 1][2*tx + dx - 1] (zero where an index is -1)
        w8 = torch.zeros(w.shape[0], 3, 8, 8, device=w.device, dtype=F32)
        w8[:, :, 1:, 1:] = w.detach().to(F32)
        w4 = w8.view(w.shape[0], 3, 4, 2, 4, 2).permute(0, 2, 4, 1, 3, 5).reshape(w.shape[0], 4, 48)   # [o][ty][tx,c,dy,dx]
        self.stem4 = PackedConv([w4.permute(0, 2, 1).reshape(w.shape[0], 48, 4, 1).contiguous()], [sd[pfx + "conv1.bias"]],
                                1, (2, 0), bn(pfx + "norm1"))
        self.stem4.algo_macs_per_pixel = w.shape[0] * 147          # the zero taps of the 8x8 embedding are not work
        self.blocks = []
        for stage, stride in ((1, 1), (2, 2), (3, 2)):
            for blk in (0, 1):
                p = f"{pfx}layer{stage}.{blk}."
                s = stride if blk == 0 else 1
                self.blocks.append(dict(
                    conv1=pc(p + "conv1", s, 1, p + "norm1"), conv2=pc(p + "conv2", 1, 1, p + "norm2"),
                    down=pc(p + "downsample.0", s, 0, p + "norm3") if s != 1 else None))
        self.head = pc(pfx + "conv2", 1, 0)

    @staticmethod
    def stem_patches(k: Kernels, images: Sequence[torch.Tensor]):
        """Operand planes of the 7x7/s2 stem (raft/extractor.py:163) in its 4-vertical-tap form: 48 channels per
        half-resolution pixel (accflow_stem_rows_planes).  They depend on the image only, so one gather serves every
        encoder that reads the same frames -> (base pointer, pitch, plane stride in elements, bytes per image)."""
        n = sum(int(im.shape[0]) for im in images)
        H, W = int(images[0].shape[-2]), int(images[0].shape[-1])
        h2, w2 = H // 2, W // 2
        pitch = Kernels.plane_pitch(48)
        patches = k.buf16("stem.rows", k.nplanes, n, h2, w2, pitch)
        stride_pl = n * h2 * w2 * pitch
        b0 = 0
        for im in images:
            assert im.dtype == F32 and im.is_contiguous() and im.shape[1] == 3
            nb = int(im.shape[0])
            L.call("accflow_stem_rows_planes", im.data_ptr(), nb, H, W,
                   patches.data_ptr() + 2 * b0 * h2 * w2 * pitch, pitch, stride_pl, k.plane_fmt, _stream())
            b0 += nb
        return patches.data_ptr(), pitch, stride_pl, 2 * h2 * w2 * pitch

    def run(self, k: Kernels, images: Sequence[torch.Tensor], tag: str, head_kwargs=None, head_out: Optional[View] = None,
            patches=None, patch_image0: int = 0) -> Optional[View]:
        """``patches``: result of ``stem_patches`` over a superset of ``images`` (same order); ``patch_image0`` is the
        index of this call's first image inside it."""
        n = sum(int(im.shape[0]) for im in images)
        H, W = int(images[0].shape[-2]), int(images[0].shape[-1])
        inst = self.norm == "instance"
        relu = L.ACT_NONE if inst else L.ACT_RELU
        h2, w2 = (H + 1) // 2, (W + 1) // 2
        x = k.view(tag + ".stem", n, h2, w2, 64)
        b0 = 0
        if k.tc:
            # x taps folded into 48 channels (operand planes), the 4 y taps served by the conv kernel's shift mode
            if patches is None:
                patches, patch_image0 = self.stem_patches(k, images), 0
            pptr, pitch, stride_pl, img_bytes = patches
            k.conv(self.stem4, [PlanesOnly(n, h2, w2, 48, pitch)], x, act=relu,
                   tc_src_planes=[(pptr + patch_image0 * img_bytes, pitch, stride_pl)], emit_planes=not inst, out_h=h2)
        else:
            for im in images:
                assert im.dtype == F32 and im.is_contiguous() and im.shape[1] == 3
                nb = int(im.shape[0])
                k.conv_smallc(im.data_ptr(), True, nb, 3, H, W, self.stem, relu, x.rows(b0, b0 + nb))
                b0 += nb
        if inst:
            k.instnorm(x, True, None, False, x)
        for bi, blk in enumerate(self.blocks):
            c1, c2, dn = blk["conv1"], blk["conv2"], blk["down"]
            oh = (x.h + 2 - 3) // c1.stride + 1
            ow = (x.w + 2 - 3) // c1.stride + 1
            y1 = k.view(f"{tag}.b{bi}.y1", n, oh, ow, c1.cout)
            y2 = k.view(f"{tag}.b{bi}.y2", n, oh, ow, c2.cout)
            if inst:
                k.conv(c1, [x], y1, emit_planes=False)                       # InstanceNorm reads fp32 and writes the planes
                k.instnorm(y1, True, None, False, y1, planes_only=True)      # only conv2 reads y1
                k.conv(c2, [y1], y2, emit_planes=False)
                res = x
                if dn is not None:
                    res = k.view(f"{tag}.b{bi}.dn", n, oh, ow, dn.cout)
                    k.conv(dn, [x], res, emit_planes=False)
                    k.instnorm(res, False, None, False, res, emit_planes=False)     # read as the fp32 residual only
                k.instnorm(y2, True, res, True, y2)
            else:
                k.conv(c1, [x], y1, act=L.ACT_RELU, planes_only=True)        # only conv2 reads y1
                res = x
                if dn is not None:
                    res = k.view(f"{tag}.b{bi}.dn", n, oh, ow, dn.cout)
                    k.conv(dn, [x], res)
                k.conv(c2, [y1], y2, act=L.ACT_RELU, residual=res, post_relu=True)
            x = y2
        if head_kwargs is not None:
            k.conv(self.head, [x], **head_kwargs)
            return None
        out = head_out if head_out is not None else k.view(tag + ".out", n, x.h, x.w, self.out_dim)
        k.conv(self.head, [x], out)
        return out


# ================================================================================ RAFT / GMA
class FlowEstimatorEngine:
    """RAFT.forward / RAFTGMA.forward (raft/raft.py:94-146, gma/gma.py:70-125) on the kernels."""

    RADIUS = 4

    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device, pfx: str = "", gma: bool = False,
                 precision: str = "fp32"):
        self.k = Kernels(device, precision)
        self.device, self.gma, self.pfx = device, gma, pfx
        self.graphs = GraphCache()
        self.repack(sd)

    def repack(self, sd):
        p, gma = self.pfx, self.gma
        sd = {n: t.detach().to(self.device) for n, t in sd.items() if n.startswith(p)}
        self.fnet = EncoderPlan(sd, p + "fnet.", "instance", 256)
        self.cnet = EncoderPlan(sd, p + "cnet.", "batch", 256)
        u = p + "update_block."

        def pc(name, pad, **kw):
            return PackedConv([sd[name + ".weight"]], [sd.get(name + ".bias")], 1, pad, **kw)

        self.convc1 = pc(u + "encoder.convc1", (0, 0))
        self.convc2 = pc(u + "encoder.convc2", (1, 1))
        self.convf1 = pc(u + "encoder.convf1", (3, 3))
        self.convf1.w = sd[u + "encoder.convf1.weight"].to(F32).permute(2, 3, 1, 0).reshape(98, 128).contiguous()
        self.convf2 = pc(u + "encoder.convf2", (1, 1))
        self.convm = pc(u + "encoder.conv", (1, 1))
        # SepConvGRU (raft/update.py:45-60) reads cat[h, inp, mf(, mf_global)].  `inp` (input channels 128..255) does
        # not change over the iterations of one pair, so the filters are split: the `inp` columns are applied once per
        # frame (gru_inp_terms) and enter every iteration as a pre-activation addend; the per-iteration convolutions
        # contract over [h, mf(, mf_global)] only (2/3 resp. 3/4 of the reference's GRU multiply-adds).
        g = u + "gru."
        self.gru, self.gru_inp = [], []
        for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
            wz, wr, wq = (sd[f"{g}conv{n}{tag}.weight"] for n in "zrq")
            rest = lambda w: torch.cat([w[:, :128], w[:, 256:]], 1)
            only = lambda w: w[:, 128:256]
            zr = PackedConv([rest(wz), rest(wr)], [sd[f"{g}convz{tag}.bias"], sd[f"{g}convr{tag}.bias"]], 1, pad)
            q = PackedConv([rest(wq)], [sd[f"{g}convq{tag}.bias"]], 1, pad)
            self.gru.append((zr, q))
            self.gru_inp.append((PackedConv([only(wz), only(wr)], [None, None], 1, pad),
                                 PackedConv([only(wq)], [None], 1, pad)))
        self.fh1 = pc(u + "flow_head.conv1", (1, 1))
        self.fh2 = pc(u + "flow_head.conv2", (1, 1))
        self.mk1 = pc(u + "mask.0", (1, 1))
        self.mk2 = pc(u + "mask.2", (0, 0), mult=0.25)
        if gma:
            self.to_qk = pc(p + "att.to_qk", (0, 0))
            self.to_v = pc(u + "aggregator.to_v", (0, 0))
            self.gamma = float(sd[u + "aggregator.gamma"].item())
            self.qk_scale = 128 ** -0.5

    # ------------------------------------------------------------------------------------
    def run_fnet(self, images: Sequence[torch.Tensor], tag: str, **patch_kw) -> View:
        """Feature encoder (InstanceNorm) on a list of (n_i,3,H,W) images -> [sum n_i, h, w, 256]."""
        return self.fnet.run(self.k, images, tag + ".fnet", **patch_kw)

    def run_cnet(self, images: Sequence[torch.Tensor], tag: str, **patch_kw):
        """Context encoder (eval BatchNorm) -> (tanh(net), relu(inp)), each [N, h, w, 128] (raft.py:115-119)."""
        k = self.k
        n = sum(int(im.shape[0]) for im in images)
        h, w = int(images[0].shape[-2]) // 8, int(images[0].shape[-1]) // 8
        hid = k.view(tag + ".h", n, h, w, 128)
        inp = k.view(tag + ".inp", n, h, w, 128)
        self.cnet.run(k, images, tag + ".cnet", head_kwargs=dict(out=hid, out2=inp, act=L.ACT_TANH, act_split=128,
                                                                 act2=L.ACT_RELU), **patch_kw)
        return hid, inp

    def gru_inp_terms(self, inp: View, tag: str):
        """conv_{z|r}(inp), conv_q(inp) of both GRU halves for a batch of context maps: the part of
        SepConvGRU's six convolutions (raft/update.py:47-58) that is constant over the refinement loop.
        -> [(zr1, q1), (zr2, q2)], fp32 [n, h, w, 256 | 128]."""
        k = self.k
        out = []
        for half, (zr_i, q_i) in enumerate(self.gru_inp):
            gzr = k.view(f"{tag}.gi_zr{half}", inp.b, inp.h, inp.w, 256)
            gq = k.view(f"{tag}.gi_q{half}", inp.b, inp.h, inp.w, 128)
            k.conv(zr_i, [inp], gzr, emit_planes=False)
            k.conv(q_i, [inp], gq, emit_planes=False)
            out.append((gzr, gq))
        return out

    def prepare(self, f1: View, f2: View, hid: View, inp: View, H: int, W: int, tag: str, gru_pre=None):
        """Correlation pyramid (+ GMA attention, + the constant GRU terms unless given) for a batch of pairs whose
        features are given."""
        B, h, w = f1.b, f1.h, f1.w
        st = dict(B=B, h=h, w=w, P=h * w, H=H, W=W, hid=hid, inp=inp)
        st["gru_pre"] = gru_pre if gru_pre is not None else self.gru_inp_terms(inp, tag)
        st["pre_mod"] = 0 if st["gru_pre"][0][0].b == B else st["gru_pre"][0][0].b     # terms shared by pairs of one frame
        st["pyr"] = self.corr_pyramid(f1, f2, tag)
        if self.gma:
            st["attn"] = self.attention(inp, tag)
        return st

    def features(self, image1: torch.Tensor, image2: torch.Tensor, tag="fe"):
        """fnet on both images + correlation pyramid + cnet (+ GMA attention)."""
        B, _, H, W = image1.shape
        assert H % 8 == 0 and W % 8 == 0 and H >= 128 and W >= 128, "H, W must be multiples of 8 and >= 128"
        fm = self.run_fnet([image1, image2], tag)
        hid, inp = self.run_cnet([image1], tag)
        return self.prepare(fm.rows(0, B), fm.rows(B, 2 * B), hid, inp, H, W, tag)

    def corr_pyramid(self, f1: View, f2: View, tag: str):
        """CorrBlock.__init__ (raft/corr.py:8-22)."""
        k = self.k
        B, h, w, D = f1.b, f1.h, f1.w, f1.c
        P = h * w
        lv = [k.buf(tag + ".pyr0", B * P, P)]
        hh, ww = h, w
        for l in range(1, 4):
            hh, ww = hh // 2, ww // 2
            lv.append(k.buf(f"{tag}.pyr{l}", B * P, hh * ww))
        if (k.tc and k.precision != "bf16x3" and w % 32 == 0 and h % 16 == 0 and (2 * w) <= 128
                and os.environ.get("ACCFLOW_FUSED_POOL", "1") != "0"):
            # level 1 comes out of the GEMM epilogue (2x2 means of the scaled volume); levels 2, 3 are pooled
            # from level 1 (a quarter of the bytes of level 0)
            k.gemm_nt(tag + ".corr", f1, f2, View(lv[0].view(B, h, w, P)), alpha=1.0 / math.sqrt(D),
                      pool_out=lv[1], pool_w=w)
            L.call("accflow_corr_pool_f32", lv[1].data_ptr(), B * P, h // 2, w // 2, lv[2].data_ptr(), lv[3].data_ptr(),
                   None, _stream())
            return lv
        k.gemm_nt(tag + ".corr", f1, f2, View(lv[0].view(B, h, w, P)), alpha=1.0 / math.sqrt(D))
        L.call("accflow_corr_pool_f32", lv[0].data_ptr(), B * P, h, w, lv[1].data_ptr(), lv[2].data_ptr(),
               lv[3].data_ptr(), _stream())
        return lv

    def attention(self, inp: View, tag: str):
        """Attention.forward (gma/modules.py:54-76), heads=1, content only -> softmax(q k^T * scale).

        Tensor-core modes: the logits are never materialised.  Pass 1 runs q k^T and keeps only per-row running
        (max, sum exp) partials (EPI_ROWSTATS); pass 2 re-runs the GEMM (K = 128: cheaper than one more pass over
        the P x P matrix) and writes exp(s - max) / sum straight into 16-bit operand planes — the A operand of the
        per-iteration aggregation.  No fp32 matrix, no softmax pass, no plane-split pass.  Returns
        ("planes", ptr, pitch, plane_stride); the exact-fp32 mode returns ("f32", tensor)."""
        k = self.k
        B, h, w = inp.b, inp.h, inp.w
        P = h * w
        qk = k.view(tag + ".qk", B, h, w, 256)
        if not k.tc:
            k.conv(self.to_qk, [inp], qk)
            attn = k.buf(tag + ".attn", B, P, P)
            k.gemm_nt(tag + ".att", qk.ch(0, 128), qk.ch(128, 256), View(attn.view(B, h, w, P)), alpha=self.qk_scale)
            k.softmax_rows(attn, B * P, P)
            return ("f32", attn)
        k.planes_ptr(qk, create=True)
        k.conv(self.to_qk, [inp], qk, planes_only=True)                  # q | k feed the GEMM only
        qp, cp, pstride = k.planes_ptr(qk, create=False)
        npl = k.nplanes
        kw = L.TcWeights(qp + 2 * 128, npl, P, 128, cp, B, pstride)      # B operand = the k half of the same planes
        qsrc = [(qp, cp, pstride)]
        parts = L.call("accflow_tc_rowstat_parts", P, k.NPROD[k.precision])
        partial = k.buf(tag + ".att_part", B * P, 2 * parts)
        stats = k.buf(tag + ".att_stats", B * P, 2)
        g = _Gemm(128, P, P)
        k.conv(g, [PlanesOnly(B, h, w, 128, cp)], View(partial.view(B, h, w, 2 * parts)), alpha=self.qk_scale,
               weight_batch_stride=1, use_affine=False, tc_b=kw, tc_src_planes=qsrc, epilogue=L.EPI_ROWSTATS, cout=P)
        L.call("accflow_softmax_stats_finalize", partial.data_ptr(), B * P, parts, stats.data_ptr(), _stream())
        Pp = (P + 7) // 8 * 8
        attn_pl = k.buf16(tag + ".attn_pl", npl, B, P, Pp)
        dst = (attn_pl.data_ptr(), Pp, B * P * Pp)
        k.conv(g, [PlanesOnly(B, h, w, 128, cp)], None, alpha=self.qk_scale, weight_batch_stride=1, use_affine=False,
               tc_b=kw, tc_src_planes=qsrc, row_stats=stats, tc_out_planes=dst)
        return ("planes",) + dst

    def aggregate(self, attn, mf: View, mfg: View, tag: str):
        """Aggregate.forward (gma/modules.py:102-115): mfg = mf + gamma * (attn @ to_v(mf)).  Tensor-core modes: the
        to_v conv writes v transposed into operand planes (EPI_STORE_T), the aggregation GEMM reads the attention
        planes of ``attention`` through TMA and adds the gamma-scaled result onto ``mf`` in its epilogue."""
        k = self.k
        B, h, w = mf.b, mf.h, mf.w
        P = h * w
        if attn[0] == "f32":
            vbuf = k.view(tag + ".v", B, h, w, 128)
            k.conv(self.to_v, [mf], vbuf)
            k.gemm_nn(tag + ".agg", View(attn[1].view(B, h, w, P)), vbuf, mfg, alpha=self.gamma, residual=mf)
            return
        _, ap, apitch, astride = attn
        npl = k.nplanes
        Pp = (P + 7) // 8 * 8
        vt = k.buf16(tag + ".vT", npl, B, 128, Pp)
        k.conv(self.to_v, [mf], None, epilogue=L.EPI_STORE_T, tc_out_planes=(vt.data_ptr(), Pp, B * 128 * Pp))
        tw = L.TcWeights(vt.data_ptr(), npl, 128, P, Pp, B, 0)
        k.conv(_Gemm(P, 128, 128), [PlanesOnly(B, h, w, P, apitch)], mfg, alpha=self.gamma, weight_batch_stride=1,
               residual=mf, use_affine=False, tc_b=tw, tc_src_planes=[(ap, apitch, astride)], planes_only=True)

    def iterate(self, st, iters: int, flow_init: Optional[torch.Tensor], tag="fe") -> torch.Tensor:
        """The GRU refinement loop + final convex upsample (raft/raft.py:121-146)."""
        k = self.k
        B, h, w, P, H, W = st["B"], st["h"], st["w"], st["P"], st["H"], st["W"]
        s = _stream
        lv = st["pyr"]
        hid, inp = st["hid"], st["inp"]
        coords = k.buf(tag + ".coords", B, P, 2)
        flow = k.buf(tag + ".flow", B, P, 2)
        corr = k.view(tag + ".corr", B, h, w, 324)
        cor1 = k.view(tag + ".cor1", B, h, w, 256)
        cf = k.view(tag + ".corflo", B, h, w, 256)
        flo1 = k.view(tag + ".flo1", B, h, w, 128)
        mf = k.view(tag + ".mf", B, h, w, 128)
        rh = k.view(tag + ".rh", B, h, w, 128)
        z = k.view(tag + ".z", B, h, w, 128)
        fh = k.view(tag + ".fh", B, h, w, 256)
        delta = k.view(tag + ".delta", B, h, w, 2)
        x_srcs = [mf]                      # `inp` enters through st["gru_pre"]
        if self.gma:
            mfg = k.view(tag + ".mfg", B, h, w, 128)
            x_srcs = [mf, mfg]
        if flow_init is not None:
            flow_init = flow_init.to(device=self.device, dtype=F32).contiguous()
            assert tuple(flow_init.shape) == (B, 2, h, w)
        L.call("accflow_coords_init_f32", None if flow_init is None else flow_init.data_ptr(), B, h, w,
               coords.data_ptr(), s())
        # Tile directions (results do not depend on them): every conv walks opposite to the producer of its input - the
        # lookup and the patch / tap-sum kernels run first-to-last - so it starts on the part still in L2.
        FWD, REV = L.TILES_FORWARD, L.TILES_REVERSE
        for _ in range(iters):
            k.corr_lookup(lv, self.RADIUS, coords, corr, flow, mf.ch(126, 128), planes_only=True)
            k.conv(self.convc1, [corr], cor1, act=L.ACT_RELU, planes_only=True, tile_order=REV)
            k.conv(self.convc2, [cor1], cf.ch(0, 192), act=L.ACT_RELU, planes_only=True, tile_order=FWD)
            k.flow_conv7(tag, flow, B, h, w, self.convf1, flo1, planes_only=True, tile_order=REV)
            k.conv(self.convf2, [flo1], cf.ch(192, 256), act=L.ACT_RELU, planes_only=True, tile_order=FWD)
            k.conv(self.convm, [cf], mf.ch(0, 126), act=L.ACT_RELU, planes_only=not self.gma, tile_order=REV)
            if self.gma:
                self.aggregate(st["attn"], mf, mfg, tag)
            for (zr, q), (pre_zr, pre_q) in zip(self.gru, st["gru_pre"]):
                k.conv(zr, [hid] + x_srcs, epilogue=L.EPI_GRU_ZR, h=hid, z=z, out2=rh, planes_only=True, pre_add=pre_zr,
                       pre_mod=st["pre_mod"], tile_order=FWD)
                k.conv(q, [rh] + x_srcs, epilogue=L.EPI_GRU_Q, h=hid, z=z, pre_add=pre_q, pre_mod=st["pre_mod"],
                       tile_order=REV)
            k.conv(self.fh1, [hid], fh, act=L.ACT_RELU, planes_only=True, tile_order=FWD)
            k.conv_smallcout(self.fh2, fh, delta, accum=coords, accum_ld=2, tile_order=REV)     # coords1 += delta_flow
        # mask head + convex upsample: only the last iteration's is observable (raft.py:139-146)
        k.conv(self.mk1, [hid], fh, act=L.ACT_RELU, planes_only=True)
        mask = k.view(tag + ".mask", B, h, w, 576)
        k.conv(self.mk2, [fh], mask)
        out = torch.empty(B, 2, H, W, device=self.device, dtype=F32)
        L.call("accflow_convex_upsample_f32", coords.data_ptr(), 2, 1, mask.ptr, mask.ld, B, h, w, out.data_ptr(), s())
        return out

    def forward(self, image1, image2, iters=12, flow_init=None, tag="fe", graph=False):
        image1 = image1.to(device=self.device, dtype=F32).contiguous()
        image2 = image2.to(device=self.device, dtype=F32).contiguous()
        if flow_init is not None:
            flow_init = flow_init.to(device=self.device, dtype=F32).contiguous()

        def eager(i1, i2, fi):
            st = self.features(i1, i2, tag)
            return self.iterate(st, iters, fi, tag)

        with torch.cuda.device(self.device):
            if not graph:
                return eager(image1, image2, flow_init)
            key = (tuple(image1.shape), iters, flow_init is not None, tag)
            return self.graphs.run(key, [image1, image2, flow_init], eager)


class _Gemm:
    """Shape carrier for per-sample GEMMs run through the conv kernel (1x1, weights given at call)."""

    def __init__(self, cin, cout, cout_pad):
        self.cin, self.cout, self.cout_pad = cin, cout, cout_pad
        self.kh = self.kw = self.stride = 1
        self.pad_h = self.pad_w = 0
        self.w = None
        self.scale = None
        self.shift = None
        self.has_bias = False


# ================================================================================ AccFlow
class AccFlowEngine:
    """AccFlow.iter / forward (networks/AccFlow_.py:157-201) on the kernels."""

    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device, gma: bool, precision: str = "fp32"):
        self.device = device
        self.ofe = FlowEstimatorEngine(sd, device, "ofe.", gma, precision)
        self.k = self.ofe.k
        self.graphs = GraphCache()
        self.repack(sd, repack_ofe=False)

    def repack(self, sd, repack_ofe=True):
        if repack_ofe:
            self.ofe.repack(sd)
        sd = {n: t.detach().to(self.device) for n, t in sd.items() if not n.startswith("ofe.")}

        def pc(name, pad, **kw):
            return PackedConv([sd[name + ".weight"]], [sd.get(name + ".bias")], 1, (pad, pad), **kw)

        self.fe1 = pc("flow_encoder.conv1", 3)
        self.fe1.w = sd["flow_encoder.conv1.weight"].to(F32).permute(2, 3, 1, 0).reshape(98, 128).contiguous()
        self.fe2 = pc("flow_encoder.conv2", 1)
        self.fe3 = pc("flow_encoder.conv3", 0)
        self.context = EncoderPlan(sd, "context.", "none", 128)
        a = "accplus."
        self.a10, self.a12 = pc(a + "conv1.0", 1), pc(a + "conv1.2", 1)
        self.a20, self.a22 = pc(a + "conv2.0", 1), pc(a + "conv2.2", 1)
        self.a24 = pc(a + "conv2.4.conv", 1, out_scale=torch.exp(sd[a + "conv2.4.scale"].to(F32).reshape(-1) * 3))
        dw = sd[a + "dconv.weight"]                                  # (128,128,3,3) -> 1x1 over 9*128
        self.dcn = PackedConv([dw.permute(0, 2, 3, 1).reshape(dw.shape[0], -1, 1, 1)], [sd[a + "dconv.bias"]], 1, (0, 0))
        self.a30, self.a32 = pc(a + "conv3.0", 1), pc(a + "conv3.2", 1)
        self.a40, self.a42, self.a44 = pc(a + "conv4.0", 1), pc(a + "conv4.2", 1), pc(a + "conv4.4", 0)
        self.bl0, self.bl2 = pc("blending.mask.0", 0), pc("blending.mask.2", 1)
        self.df0, self.df2 = pc("flow_decoder.flow.0", 1), pc("flow_decoder.flow.2", 1)
        self.dm0, self.dm2 = pc("flow_decoder.mask.0", 1), pc("flow_decoder.mask.2", 0)

    def iter(self, I1, I2, In, F2n: Optional[torch.Tensor], iters=12):
        """Returns (out_small NCHW (b,2,h,w), out NCHW (b,2,H,W)); F2n is NCHW (b,2,h,w) or None."""
        dev = self.device
        I1, I2, In = (t.to(device=dev, dtype=F32).contiguous() for t in (I1, I2, In))
        b = I1.shape[0]
        with torch.cuda.device(dev):
            if F2n is None:
                flows = self.ofe.forward(torch.cat([I1, I1, I2]), torch.cat([I2, In, In]), iters, tag="ofe3")
            else:
                flows = self.ofe.forward(torch.cat([I1, I1]), torch.cat([I2, In]), iters, tag="ofe2")
            ctx = self.context.run(self.k, [I1, I2, In], "acc.ctx")
            return self._accumulate(flows, b, ctx.rows(0, b), ctx.rows(b, 2 * b), ctx.rows(2 * b, 3 * b), F2n)

    def _accumulate(self, flows: torch.Tensor, b: int, c1: View, c2: View, cn: View, F2n: Optional[torch.Tensor]):
        """Everything of AccFlow.iter after the ofe call (AccFlow_.py:185-201), given the context features."""
        k, s = self.k, _stream
        dev = self.device
        H, W = int(flows.shape[-2]), int(flows.shape[-1])
        assert H % 8 == 0 and W % 8 == 0
        h, w = H // 8, W // 8
        P = h * w
        npair = int(flows.shape[0]) // b
        lr = k.buf("acc.lr", npair * b, P, 2)                       # [dflow | flow_ini | (F2n)]
        L.call("accflow_downflow8_f32", flows.data_ptr(), npair * b, H, W, lr.data_ptr(), s())
        fin = k.buf("acc.fin", 3 * b, P, 2)                          # encoder order: flow_ini, dflow, F2n
        fin[0:b].copy_(lr[b:2 * b])
        fin[b:2 * b].copy_(lr[0:b])
        if F2n is None:
            fin[2 * b:].copy_(lr[2 * b:])
        else:
            fin[2 * b:].copy_(F2n.to(device=dev, dtype=F32).permute(0, 2, 3, 1).reshape(b, P, 2))
        dflow, flow_ini = fin[b:2 * b], fin[0:b]
        # FlowEncoder (AccFlow_.py:56-65)
        e1 = k.view("acc.e1", 3 * b, h, w, 128)
        e2 = k.view("acc.e2", 3 * b, h, w, 256)
        enc = k.view("acc.enc", 3 * b, h, w, 128)
        k.flow_conv7("acc.fe", fin, 3 * b, h, w, self.fe1, e1)
        k.conv(self.fe2, [e1], e2, act=L.ACT_RELU)
        k.conv(self.fe3, [e2], enc)
        f_ini, df, f = enc.rows(0, b), enc.rows(b, 2 * b), enc.rows(2 * b, 3 * b)
        # occlusion + error maps (getOcc, AccFlow_.py:127-135,194,197)
        occ = k.view("acc.occ", b, h, w, 1)
        emap = k.view("acc.emap", b, h, w, 128)
        L.call("accflow_warp_occ_f32", c1.ptr, c1.ld, c2.ptr, c2.ld, dflow.data_ptr(), b, h, w, 128, occ.ptr, 1,
               None, 0, s())
        L.call("accflow_warp_occ_f32", c1.ptr, c1.ld, cn.ptr, cn.ld, flow_ini.data_ptr(), b, h, w, 128, None, 0,
               emap.ptr, emap.ld, s())
        k.wrote(occ)
        k.wrote(emap)
        # AccPlus (AccFlow_.py:97-109)
        t256 = k.view("acc.t256", b, h, w, 256)
        x1 = k.view("acc.x1", b, h, w, 128)
        x2 = k.view("acc.x2", b, h, w, 128)
        om = k.view("acc.om", b, h, w, 28)
        col = k.buf("acc.col", b, P, 9 * 128)
        fdc = k.view("acc.fdc", b, h, w, 128)
        k.conv(self.a10, [df, f, occ], t256, act=L.ACT_RELU)
        k.conv(self.a12, [t256], x1)
        k.conv(self.a20, [x1, c1], t256, act=L.ACT_RELU)
        k.conv(self.a22, [t256], x2, act=L.ACT_RELU)
        k.conv(self.a24, [x2], om.ch(0, 27))
        L.call("accflow_deform_gather_f32", f.ptr, f.ld, om.ptr, om.ld, b, h, w, 128, col.data_ptr(), s())
        colv = View(col.view(b, h, w, 9 * 128))
        k.wrote(colv)
        k.conv(self.dcn, [colv], fdc)
        k.conv(self.a30, [fdc, df, occ], t256, act=L.ACT_RELU)
        k.conv(self.a32, [t256], x1)
        k.conv(self.a40, [x1, c1, fdc, df], t256, act=L.ACT_RELU)
        k.conv(self.a42, [t256], x2, act=L.ACT_RELU)
        f_acc = k.view("acc.facc", b, h, w, 128)
        k.conv(self.a44, [x2], f_acc)
        # Blending (AccFlow_.py:122-124)
        m = k.view("acc.m", b, h, w, 1)
        k.conv(self.bl0, [emap], t256, act=L.ACT_RELU)
        k.conv_smallcout(self.bl2, t256, m, act=L.ACT_SIGMOID)
        fuse = k.view("acc.fuse", b, h, w, 128)
        L.call("accflow_blend_f32", f_ini.ptr, f_acc.ptr, m.ptr, 1, b * P, 128, fuse.ptr, s())
        k.wrote(fuse)
        # FlowDecoder (AccFlow_.py:40-45)
        small = torch.empty(b, h, w, 2, device=dev, dtype=F32)
        k.conv(self.df0, [fuse], t256, act=L.ACT_RELU)
        k.conv_smallcout(self.df2, t256, View(small))
        mask = k.view("acc.mask", b, h, w, 576)
        k.conv(self.dm0, [fuse], t256, act=L.ACT_RELU)
        k.conv(self.dm2, [t256], mask)
        out = torch.empty(b, 2, H, W, device=dev, dtype=F32)
        L.call("accflow_convex_upsample_f32", small.data_ptr(), 2, 0, mask.ptr, mask.ld, b, h, w, out.data_ptr(), s())
        self.last_dflow = dflow.view(b, h, w, 2).permute(0, 3, 1, 2).contiguous()     # F(i -> i-1) at 1/8 (warm start)
        return small.permute(0, 3, 1, 2).contiguous(), out

    def forward(self, images: List[torch.Tensor], iters=12, graph=False, warm_start=False,
                warm_iters: Optional[int] = None) -> List[torch.Tensor]:
        """AccFlow.forward (AccFlow_.py:157-175) with every encoder evaluated once per distinct frame.

        The reference re-runs fnet / cnet / context on the same frames at every accumulation step
        (22 / 11 / 15 image passes per 7-frame clip for 7 / 6 / 7 distinct frames);
        all three encoders are per-sample functions, so the cached features are identical.

        ``warm_start`` (the reference README's open TODO "Add warmstart mode"; the estimators already take
        ``flow_init``, raft/raft.py:123-124): from the second accumulation step on, the pair (i -> i-1) starts from
        the previous step's F(i-1 -> i-2) and the pair (i -> 0) from the previous accumulated flow F(i-1 -> 0), both
        at 1/8 resolution, and may run ``warm_iters`` (<= iters) refinement iterations.
        """
        images = [t.to(device=self.device, dtype=F32).contiguous() for t in images]

        def eager(*imgs):
            k, ofe = self.k, self.ofe
            n, b = len(imgs), int(imgs[0].shape[0])
            H, W = int(imgs[0].shape[-2]), int(imgs[0].shape[-1])
            assert H % 8 == 0 and W % 8 == 0 and H >= 128 and W >= 128, "H, W must be multiples of 8 and >= 128"
            # the three encoders share one stem im2col of the frames (tensor-core modes)
            pk = dict(patches=EncoderPlan.stem_patches(k, imgs)) if k.tc else {}
            fm = ofe.run_fnet(list(imgs), "clip", **pk)                    # frame f -> rows [f*b, (f+1)*b)
            hid_all, inp_all = ofe.run_cnet(list(imgs[1:]), "clip", patch_image0=b, **pk)  # frame f (>=1) -> rows [(f-1)*b, f*b)
            ctx = self.context.run(k, list(imgs), "clip.ctx", **pk)
            gi_all = ofe.gru_inp_terms(inp_all, "clip")                   # constant GRU terms, once per frame
            h, w = fm.h, fm.w

            def gather(src: View, frames, name):
                dst = k.view(name, len(frames) * b, h, w, src.c)
                for j, f in enumerate(frames):
                    dst.t[j * b:(j + 1) * b].copy_(src.t[f * b:(f + 1) * b])
                k.wrote(dst)
                return dst

            flow, outs = None, []
            for i in range(2, n):
                pairs = [(i, i - 1), (i, 0), (i - 1, 0)] if flow is None else [(i, i - 1), (i, 0)]
                tag = f"ofe{len(pairs)}"
                f1 = gather(fm, [p[0] for p in pairs], tag + ".f1")
                f2 = gather(fm, [p[1] for p in pairs], tag + ".f2")
                hid = gather(hid_all, [p[0] - 1 for p in pairs], tag + ".hid")
                inp = gather(inp_all, [p[0] - 1 for p in pairs], tag + ".inpg") if ofe.gma else None   # attention only
                if len({p[0] for p in pairs}) == 1:       # (i, i-1), (i, 0): one term per clip, read by both pairs (pre_mod)
                    pre = [tuple(t.rows((i - 1) * b, i * b) for t in pr) for pr in gi_all]
                else:
                    pre = [tuple(gather(t, [p[0] - 1 for p in pairs], f"{tag}.gi{hf}{j}") for j, t in enumerate(pr))
                           for hf, pr in enumerate(gi_all)]
                st = ofe.prepare(f1, f2, hid, inp, H, W, tag, gru_pre=pre)
                if warm_start and flow is not None:
                    flows = ofe.iterate(st, warm_iters or iters, torch.cat([self.last_dflow, flow]), tag)
                else:
                    flows = ofe.iterate(st, iters, None, tag)
                flow, up = self._accumulate(flows, b, ctx.rows(i * b, (i + 1) * b), ctx.rows((i - 1) * b, i * b),
                                            ctx.rows(0, b), flow)
                outs.append(up)
            return outs

        with torch.cuda.device(self.device):
            if not graph:
                return eager(*images)
            key = (tuple(images[0].shape), len(images), iters, bool(warm_start), warm_iters)
            return self.graphs.run(key, images, eager)
