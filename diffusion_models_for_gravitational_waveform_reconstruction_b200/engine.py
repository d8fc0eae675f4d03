"""Host-side orchestration of the sm_100a kernels for UNet1D.forward and the DDIM/DDPM chain.

Everything numerical happens in libgwb200.so (csrc/*.cu) through the C-ABI; torch provides device memory,
streams and CUDA-graph capture only.  Layer structure follows models.py:195-231 of the reference.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _cabi
from ._cabi import GW_BF16, GW_F32, ConvTcShape, StepParams, check, ptr

Tensor = torch.Tensor


@dataclass
class ModelSpec:
    in_ch: int = 1
    base_ch: int = 64
    time_dim: int = 128
    depth: int = 3
    kernel: int = 3
    max_time: float = 999.0
    cond_in_ch: int = 0
    use_selfcond: bool = False

    @property
    def chs(self) -> List[int]:
        return [self.base_ch * (2 ** i) for i in range(self.depth)]

    @property
    def layer_channels(self) -> List[int]:
        """enc0..encD-1, mid, dec0..decD-1"""
        c = self.chs
        return c + [c[-1]] + list(reversed(c))

    @property
    def film_dim(self) -> int:
        return 2 * sum(self.layer_channels)

    def film_offsets(self) -> List[int]:
        offs, o = [], 0
        for c in self.layer_channels:
            offs.append(o)
            o += 2 * c
        return offs

    def level_lengths(self, L: int) -> List[int]:
        out = [L]
        for _ in range(self.depth):
            out.append(out[-1] // 2)
        return out

    def layer_names(self) -> List[str]:
        d = self.depth
        return [f"encoders.{i}" for i in range(d)] + ["mid"] + [f"decoders.{i}" for i in range(d)]

    def cond_names(self) -> List[str]:
        d = self.depth
        return [f"cond_enc.{i}" for i in range(d)] + ["cond_mid"] + [f"cond_dec.{i}" for i in range(d)]

    def tproj_names(self) -> List[str]:
        d = self.depth
        return [f"tproj_enc.{i}.1" for i in range(d)] + ["tproj_mid.1"] + [f"tproj_dec.{i}.1" for i in range(d)]

    def conv_flops(self, L: int) -> float:
        """Algorithmic FLOPs of one forward of one sample (2*Cin*Cout*K*L over every Conv1d)."""
        Ls = self.level_lengths(L)
        fl, cin = 0.0, self.in_ch
        for i, c in enumerate(self.chs):
            fl += 2.0 * cin * c * self.kernel * Ls[i] + 2.0 * self.cond_in_ch * c * Ls[i]
            cin = c
        fl += 2.0 * cin * cin * self.kernel * Ls[-1] + 2.0 * self.cond_in_ch * cin * Ls[-1]
        prev = cin
        for i, c in enumerate(reversed(self.chs)):
            Li = Ls[self.depth - 1 - i]
            fl += 2.0 * (prev + c) * c * self.kernel * Li + 2.0 * self.cond_in_ch * c * Li
            prev = c
        return fl + 2.0 * (prev + 1) * self.kernel * L


class ParamLayout:
    """Canonical flat ordering of the 60 parameter tensors: every tproj weight first (so the concatenated FiLM
    projection [film_dim, base_ch] is one contiguous block), then every tproj bias, then the rest in state_dict order.
    Offsets are multiples of 4 floats so every view is 16-byte aligned; the padding stays zero in all flat buffers."""

    def __init__(self, spec: ModelSpec, shapes: Dict[str, tuple]):
        tw = [n + ".weight" for n in spec.tproj_names()]
        tb = [n + ".bias" for n in spec.tproj_names()]
        rest = [k for k in shapes if k not in tw and k not in tb]
        self.names = tw + tb + rest
        self.shapes = {k: tuple(shapes[k]) for k in self.names}
        self.offsets: Dict[str, int] = {}
        o = 0
        for k in self.names:
            self.offsets[k] = o
            n = 1
            for d_ in self.shapes[k]:
                n *= d_
            # the tproj blocks must stay gap-free (their sizes are multiples of 4 already)
            o += (n + 3) // 4 * 4
        self.total = o
        self.w2 = (self.offsets[tw[0]], spec.film_dim * spec.base_ch)
        self.b2 = (self.offsets[tb[0]], spec.film_dim)

    def views(self, flat: Tensor) -> Dict[str, Tensor]:
        out = {}
        for k in self.names:
            n = 1
            for d_ in self.shapes[k]:
                n *= d_
            out[k] = flat[self.offsets[k]: self.offsets[k] + n].view(self.shapes[k])
        return out


class _Workspace:
    """Activation buffers for one (batch, length).  Channels-last [B, L_lvl, C] in the engine's storage dtype."""

    def __init__(self, spec: ModelSpec, B: int, L: int, tdtype: torch.dtype, device, keep_raw: bool):
        self.B, self.L = B, L
        d = spec.depth
        Ls = spec.level_lengths(L)
        self.Ls = Ls
        chs = spec.chs
        e = lambda *s: torch.empty(*s, device=device, dtype=tdtype)
        lay_len = [Ls[i] for i in range(d)] + [Ls[d]] + [Ls[d - 1 - i] for i in range(d)]
        lay_ch = spec.layer_channels
        self.lay_len = lay_len
        if keep_raw:
            self.raw = [e(B, lay_len[i], lay_ch[i]) for i in range(2 * d + 1)]
        else:
            big = e(B * max(lay_len[i] * lay_ch[i] for i in range(2 * d + 1)))
            self.raw = [big[: B * lay_len[i] * lay_ch[i]].view(B, lay_len[i], lay_ch[i]) for i in range(2 * d + 1)]
        self.out = [e(B, lay_len[i], lay_ch[i]) for i in range(2 * d + 1)]
        self.pooled = [e(B, Ls[i + 1], chs[i]) for i in range(d)]
        n_part_max = (L + 63) // 64 + 2
        self.part = torch.empty(B, n_part_max, 8, 2, device=device, dtype=torch.float32)
        self.stats = [torch.empty(B, 8, 2, device=device, dtype=torch.float32) for _ in range(2 * d + 1)] if keep_raw else None
        self.cond = [torch.empty(B, Ls[j], max(spec.cond_in_ch, 1), device=device, dtype=torch.float32)
                     for j in range(d + 1)] if spec.cond_in_ch > 0 else None
        self.sync: Optional[Tensor] = None          # exchange buffer of the first block (zeroed once, see UNetEngine.workspace)
        self.syncs: List[Optional[Tensor]] = [None] * (2 * d + 1)   # gw_conv_gn: one buffer per layer (packets, epochs, chain flags)
        self.chain_prev = None                      # (sync buffer, CTAs per sample) of the layer that just ran with chain signalling
        self.coef0: Optional[Tensor] = None         # gw_conv_in_direct: per-(sample, channel) epilogue coefficients of block 0
        self.dots: Optional[Tensor] = None          # [B, L, 4] head dot products left by the last decoder's fused kernel
        self.head_fused = False                     # set by UNetEngine.body(): ws.dots is current, ws.out[-1] was not written


class UNetEngine:
    """Runs the denoiser forward on the GPU.  `params` maps reference state_dict keys to CUDA fp32 tensors
    (the engine keeps references, not copies, so in-place optimiser updates are seen after `refresh()`)."""

    def __init__(self, params: Dict[str, Tensor], spec: ModelSpec, dtype: str = "bf16", conv_impl: str = "auto",
                 tc_variant: int = 6):
        self.lib = _cabi.load()
        self.spec = spec
        assert dtype in ("bf16", "fp32")
        self.dtype = dtype
        self.gw_dtype = GW_BF16 if dtype == "bf16" else GW_F32
        self.tdtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        if conv_impl == "auto":
            conv_impl = "tc" if dtype == "bf16" else "simt"
        if conv_impl == "tc" and dtype != "bf16":
            raise ValueError("the tcgen05 conv path stores bf16 activations; use dtype='bf16'")
        self.conv_impl = conv_impl
        self.tc_variant = tc_variant
        # Non-default architectures (models.py:78-88 takes base_ch and kernel; the training CLI exposes --base_ch, train.py:641): every layer runs the shape-generic
        # CUDA-core kernels of csrc/generic.cu (fp32 math, fp32 or bf16 storage) instead of the tcgen05 / streaming kernels,
        # which are laid out for 64-channel rows and three taps.
        self.generic = spec.base_ch % 64 != 0 or spec.kernel != 3
        if self.generic:
            b = spec.base_ch
            if b < 4 or (b & (b - 1)) != 0 or b > 1024:
                raise ValueError(f"base_ch={b}: the time-MLP kernels need a power of two in [4, 1024] (or a multiple of 64 with "
                                 "kernel=3)")
            if spec.kernel not in (1, 3, 5, 7):
                raise ValueError(f"kernel={spec.kernel}: odd kernel sizes up to 7 ('same' padding needs an odd size)")
            if spec.cond_in_ch > 8:
                raise ValueError("cond_in_ch > 8")
            self.conv_impl = "simt"
        self.p = params
        dev = params["final.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("UNetEngine needs CUDA tensors: there is no CPU fallback for this path")
        self.device = dev
        self._ws: Dict[tuple, _Workspace] = {}
        self._packed: Dict[tuple, Tensor] = {}
        self.launches = 0
        self._chain_serial: Optional[Tensor] = None
        # inference option: gw_conv_in_block (stats pass + recompute/apply pass, no raw tensor) instead of gw_conv_in +
        # gw_gn_apply.  Measured on B200 at B=256, L=4096: 166 us vs 70 + 72 us -- the CUDA-core conv is issue-bound, so
        # recomputing it costs more than the 268 MB of HBM traffic it saves; kept (parity-tested) but off by default.
        self.fuse_first_block = False
        self.stream_gn = True                       # bf16: bulk-copy streaming GroupNorm kernels (stream_gn.cu)
        # bf16/tcgen05: conv + GroupNorm + SiLU + cond + FiLM (+ pool) of a block in ONE kernel (conv_gn.cuh) wherever the
        # layer shape allows it; the conv output then never makes the HBM round trip between gw_conv_tc and gw_gn_apply
        self.fuse_gn = True
        # training (keep_raw) forwards also have to write the conv output for the backward pass and carry five cond channels;
        # measured on B200 the fused kernel then loses its edge (962 us vs ~900 us of conv + GroupNorm launches per step at
        # B=256, L=4096, in_ch=7), so it stays an inference default; parity-tested for both
        self.fuse_gn_train = os.environ.get("GWB200_FUSE_GN_TRAIN", "0") not in ("0", "")
        # inference: the last decoder's fused kernel also forms the three head-conv dot products per position (gw_conv_gn2), so
        # gw_final_step runs on 16 B per position and the [B, L, 64] activation is neither written nor read back
        self.fuse_head = True
        # sampler: chain the fused layer kernels (gw_conv_gn3): each layer's launch overlaps the tail of the previous one and
        # orders itself per sample through completion flags instead of a grid-wide dependency.  Measured on B200 (B=256, L=4096;
        # profiles/r02_chain_overlap.md): the six layer kernels of an eager reverse step finish 19 us earlier (706 vs 725 us: the
        # 6 us launch gaps disappear, the overlap is real), but inside the CUDA graph the chain is power-capped -- the SM clock
        # drops from 1837 to 1788 MHz and the throughput is unchanged (287.7 vs 289.3 waveforms/s) -- so it stays an option
        # (GWB200_CHAIN=1), parity-tested in both modes.  Small batches are a different regime: at B=8 the chain is latency-bound
        # at full clock (64 of 148 SMs busy, ~14 us per layer kernel) and chaining is worth +14 % (80.4 vs 70.5 waveforms/s,
        # profiles/r02_bench_ddpm1000_B8_chain.json), so `None` = automatic: on when a layer's samples fit in two rounds of CTA
        # groups (batch <= 36), off above.  GWB200_CHAIN=0 / 1 or setting the attribute forces it.
        env = os.environ.get("GWB200_CHAIN", "")
        self.chain_layers: Optional[bool] = None if env == "" else (env != "0")
        self.chain_auto_batch = 36
        # inference: first block in one pass with ANALYTIC GroupNorm statistics (conv_in_direct.cu) instead of the exchange-based
        # conv_in_gn kernel.  Measured on B200 (B=256, L=4096, in_ch=3; profiles/r02_first_block.md): moments 25 us + one-pass
        # kernel 79 us vs 107 us for conv_in_gn -- a wash so far, so the parity-tested kernel stays an option
        self.direct_first = os.environ.get("GWB200_DIRECT_FIRST", "0") not in ("0", "")
        self._fuse_ok: Dict[tuple, bool] = {}
        self.flat: Optional[Tensor] = None          # set by bind_flat(): params are views of one ParamLayout buffer
        self.layout: Optional[ParamLayout] = None
        self.generation = 0                         # bumped by refresh(): weight-derived caches (SamplerPlan.film) follow it
        self.ptr_generation = 0                     # bumped when a parameter tensor MOVED: captured graphs hold raw pointers
        self._ptr_sig = None
        self.refresh()

    # ------------------------------------------------------------------ weights
    def refresh(self) -> None:
        """Re-derive kernel-side weight layouts from the current parameter values."""
        sp, p = self.spec, self.p
        if self.flat is not None:
            lo = self.layout
            self.film_w2 = self.flat[lo.w2[0]: lo.w2[0] + lo.w2[1]]
            self.film_b2 = self.flat[lo.b2[0]: lo.b2[0] + lo.b2[1]]
        else:
            self.film_w2 = torch.cat([p[n + ".weight"] for n in sp.tproj_names()], dim=0).contiguous()
            self.film_b2 = torch.cat([p[n + ".bias"] for n in sp.tproj_names()], dim=0).contiguous()
        self.wf = p["final.weight"].reshape(-1)                   # [(C+1)*3], index c*3+k
        for key in list(self._packed):
            self._pack_tc(key)
        self.generation += 1
        sig = tuple(t.data_ptr() for t in p.values())
        if sig != self._ptr_sig:                    # e.g. FusedTrainStep re-pointed the module's parameters at its flat buffer
            self._ptr_sig = sig
            self.ptr_generation += 1

    def bind_flat(self, flat: Tensor, layout: ParamLayout) -> None:
        """Use `flat` (fp32, ParamLayout order) as the parameter storage: in-place optimiser updates of the flat buffer
        are seen by every kernel without copies (only the packed bf16 conv weights need `refresh()`)."""
        self.flat, self.layout = flat, layout
        self.p = layout.views(flat)
        self.refresh()

    def _shape(self, li: int, B: int, L: int, L0: int) -> ConvTcShape:
        sp = self.spec
        d = sp.depth
        lc = sp.layer_channels
        if li <= d:
            return ConvTcShape(1, 0, B, L, lc[li - 1], L, 0, lc[li])
        i = li - d - 1
        return ConvTcShape(2, 1, B, L, lc[li - 1], L0, sp.chs[d - 1 - i], lc[li])

    def _pack_tc(self, key) -> Tensor:
        li, L, L0 = key
        shp = self._shape(li, 1, L, L0)
        n = self.lib.gw_conv_tc_packed_elems(C.byref(shp))
        if n < 0:
            raise RuntimeError("gw_conv_tc_packed_elems: " + self.lib.gw_last_error().decode())
        buf = self._packed.get(key)
        if buf is None or buf.numel() != n:
            buf = torch.empty(n, device=self.device, dtype=torch.bfloat16)
            self._packed[key] = buf
        w = self.p[self.spec.layer_names()[li] + ".0.weight"]
        check(self.lib.gw_conv_tc_pack(C.byref(shp), ptr(w), ptr(buf), _cabi.stream_ptr()), "conv_tc_pack")
        return buf

    def tc_supported(self, li: int, L: int, L0: int) -> bool:
        sp = self.spec
        if self.conv_impl != "tc" or li == 0:
            return False
        lc = sp.layer_channels
        cout = lc[li]
        if cout > 256 or (cout & (cout - 1)) != 0:
            return False
        if li > sp.depth and (L % 2 != 0 or L0 * 2 != L):
            return False
        shp = self._shape(li, 1, L, L0)
        return self.lib.gw_conv_tc_packed_elems(C.byref(shp)) > 0

    # ------------------------------------------------------------------ workspace
    def workspace(self, B: int, L: int, keep_raw: bool = False) -> _Workspace:
        key = (B, L, keep_raw)
        ws = self._ws.get(key)
        if ws is None:
            ws = _Workspace(self.spec, B, L, self.tdtype, self.device, keep_raw)
            if self.dtype == "bf16":
                # exchange buffers are zeroed ONCE, outside any graph capture: epochs / chain counters only ever advance
                nb = self.lib.gw_conv_gn_sync_bytes(B)
                ws.sync = torch.zeros(nb, device=self.device, dtype=torch.uint8)
                ws.syncs = [torch.zeros(nb, device=self.device, dtype=torch.uint8) for _ in range(2 * self.spec.depth + 1)]
            if self.generic and ws.stats is None:
                ws.stats = [torch.empty(B, 8, 2, device=self.device, dtype=torch.float32)
                            for _ in range(2 * self.spec.depth + 1)]
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ kernels
    def film_vectors(self, t: Tensor, out: Optional[Tensor] = None, aux: Optional[Tensor] = None) -> Tensor:
        """FiLM rows [n, film_dim] for integer timesteps t; `aux` [n, time_dim + 3*base_ch] keeps the MLP activations."""
        sp = self.spec
        t = t.to(device=self.device, dtype=torch.int64).contiguous()
        if out is None:
            out = torch.empty(t.numel(), sp.film_dim, device=self.device, dtype=torch.float32)
        check(self.lib.gw_film_vectors(ptr(t), t.numel(), sp.time_dim, sp.max_time, ptr(self.p["time_mlp.1.weight"]),
                                       ptr(self.p["time_mlp.1.bias"]), ptr(self.film_w2), ptr(self.film_b2),
                                       sp.base_ch, sp.film_dim, ptr(out), ptr(aux), _cabi.stream_ptr()), "film_vectors")
        self.launches += 1
        return out

    def cond_pyramid(self, ws: _Workspace, net: Tensor) -> None:
        sp = self.spec
        if sp.cond_in_ch == 0:
            return
        B, Cx, L = net.shape
        n = sp.depth + 1
        lens = (C.c_int * n)(*ws.Ls)
        outs = (C.c_void_p * n)(*[ptr(c) for c in ws.cond])
        check(self.lib.gw_cond_pyramid(ptr(net), B, Cx, L, sp.cond_in_ch, n, lens, outs, _cabi.stream_ptr()), "cond_pyramid")
        self.launches += 1

    def _conv(self, li: int, src0: Tensor, src1: Optional[Tensor], raw: Tensor, part: Tensor) -> int:
        """Conv of layer li (>=1).  Returns n_part."""
        sp = self.spec
        B, L, Cout = raw.shape
        L0, C0 = src0.shape[1], src0.shape[2]
        up = src1 is not None
        name = sp.layer_names()[li]
        st = _cabi.stream_ptr()
        if self.tc_supported(li, L, L0):
            key = (li, L, L0)
            packed = self._packed.get(key)
            if packed is None:
                packed = self._pack_tc(key)
                self.launches += 1
            shp = self._shape(li, B, L, L0)
            check(self.lib.gw_conv_tc(C.byref(shp), ptr(src0), ptr(src1), ptr(packed), ptr(self.p[name + ".0.bias"]),
                                      ptr(raw), ptr(part), self.tc_variant, st), f"conv_tc[{name}]")
            self.launches += 1
            return self.lib.gw_conv_tc_n_part(C.byref(shp))
        C1 = src1.shape[2] if up else 0
        check(self.lib.gw_conv3_simt(ptr(src0), C0, L0, 1 if up else 0, ptr(src1), C1, B, L,
                                     ptr(self.p[name + ".0.weight"]), ptr(self.p[name + ".0.bias"]), Cout, ptr(raw),
                                     self.gw_dtype, ptr(part), st), f"conv3_simt[{name}]")
        self.launches += 1
        return (L + 63) // 64

    def _gn(self, li: int, ws: _Workspace, n_part: int, film: Tensor, film_b_stride: int, film_step_stride: int,
            step_ptr: Optional[Tensor], pooled: Optional[Tensor], lvl: int) -> None:
        sp = self.spec
        name = sp.layer_names()[li]
        raw, out = ws.raw[li], ws.out[li]
        B, L, Cc_ = raw.shape
        Cc = sp.cond_in_ch
        cname = sp.cond_names()[li]
        if self.stream_gn and self.dtype == "bf16" and L % 4 == 0 and Cc_ in (64, 128, 256):
            check(self.lib.gw_gn_apply_stream(ptr(raw), ptr(ws.part), n_part, B, L, Cc_, ptr(self.p[name + ".1.weight"]),
                                              ptr(self.p[name + ".1.bias"]), ptr(ws.cond[lvl]) if Cc > 0 else None, Cc,
                                              ptr(self.p[cname + ".weight"]) if Cc > 0 else None,
                                              ptr(self.p[cname + ".bias"]) if Cc > 0 else None, ptr(film),
                                              sp.film_offsets()[li], film_b_stride, film_step_stride, ptr(step_ptr), ptr(out),
                                              ptr(pooled), ptr(ws.stats[li]) if ws.stats is not None else None,
                                              _cabi.stream_ptr()), f"gn_apply_stream[{name}]")
            self.launches += 1
            return
        check(self.lib.gw_gn_apply(ptr(raw), ptr(ws.part), n_part, B, L, Cc_, ptr(self.p[name + ".1.weight"]),
                                   ptr(self.p[name + ".1.bias"]), ptr(ws.cond[lvl]) if Cc > 0 else None, Cc,
                                   ptr(self.p[cname + ".weight"]) if Cc > 0 else None,
                                   ptr(self.p[cname + ".bias"]) if Cc > 0 else None, ptr(film), sp.film_offsets()[li],
                                   film_b_stride, film_step_stride, ptr(step_ptr), ptr(out), ptr(pooled),
                                   ptr(ws.stats[li]) if ws.stats is not None else None, self.gw_dtype,
                                   _cabi.stream_ptr()), f"gn_apply[{name}]")
        self.launches += 1

    def _block(self, li: int, ws: _Workspace, src0: Tensor, src1: Optional[Tensor], film: Tensor, film_b_stride: int,
               film_step_stride: int, step_ptr: Optional[Tensor], pooled: Optional[Tensor], lvl: int, head: bool = False) -> bool:
        """One conv block (models.py:160-173 + cond bias + FiLM (+ pool)) of layer li >= 1 into ws.out[li]."""
        sp = self.spec
        raw = ws.raw[li]
        B, L, Cout = raw.shape
        L0 = src0.shape[1]
        Cc = sp.cond_in_ch
        fused = False
        if (self.fuse_gn and (ws.stats is None or self.fuse_gn_train) and self.dtype == "bf16"
                and self.tc_supported(li, L, L0)):
            fkey = (li, L, L0, pooled is not None)
            fused = self._fuse_ok.get(fkey)
            if fused is None:
                shp1 = self._shape(li, 1, L, L0)
                fused = self.lib.gw_conv_gn_group(C.byref(shp1), Cc, 1 if pooled is not None else 0)
                self._fuse_ok[fkey] = fused
        if not fused:
            ws.chain_prev = None
            n_part = self._conv(li, src0, src1, raw, ws.part)
            self._gn(li, ws, n_part, film, film_b_stride, film_step_stride, step_ptr, pooled, lvl)
            return False
        if ws.syncs[li] is None:
            ws.syncs[li] = torch.zeros(self.lib.gw_conv_gn_sync_bytes(B), device=self.device, dtype=torch.uint8)
        key = (li, L, L0)
        packed = self._packed.get(key)
        if packed is None:
            packed = self._pack_tc(key)
            self.launches += 1
        name, cname = sp.layer_names()[li], sp.cond_names()[li]
        shp = self._shape(li, B, L, L0)
        keep = ws.stats is not None
        head = head and not keep and Cout == 64 and src1 is not None and pooled is None
        if head and ws.dots is None:
            ws.dots = torch.empty(B, L, 4, device=self.device, dtype=torch.float32)
        chain = self.chain_layers if self.chain_layers is not None else B <= self.chain_auto_batch
        serial = self._chain_serial if (chain and not keep and step_ptr is not None) else None
        prev, prev_g = ws.chain_prev if (serial is not None and ws.chain_prev is not None and ws.chain_prev[1] <= 8
                                         and os.environ.get("GWB200_CHAIN", "0") != "2") else (None, 0)
        check(self.lib.gw_conv_gn3(C.byref(shp), ptr(src0), ptr(src1), ptr(packed), ptr(self.p[name + ".0.bias"]),
                                   ptr(self.p[name + ".1.weight"]), ptr(self.p[name + ".1.bias"]),
                                   ptr(ws.cond[lvl]) if Cc > 0 else None, Cc,
                                   ptr(self.p[cname + ".weight"]) if Cc > 0 else None,
                                   ptr(self.p[cname + ".bias"]) if Cc > 0 else None, ptr(film), sp.film_offsets()[li],
                                   film_b_stride, film_step_stride, ptr(step_ptr), None if head else ptr(ws.out[li]), ptr(pooled),
                                   ptr(raw) if keep else None, ptr(ws.stats[li]) if keep else None, ptr(ws.syncs[li]),
                                   ptr(self.wf) if head else None, ptr(ws.dots) if head else None, ptr(prev), prev_g, ptr(serial),
                                   _cabi.stream_ptr()), f"conv_gn[{name}]")
        self.launches += 1
        # the next layer may chain to this launch (the head flavour writes its dots with plain stores: no flags)
        ws.chain_prev = (ws.syncs[li], int(fused)) if (serial is not None and not head) else None
        return head

    def body(self, ws: _Workspace, net_a: Tensor, net_b: Optional[Tensor], step_ptr: Optional[Tensor], film: Tensor,
             film_b_stride: int, film_step_stride: int, chain_serial: Optional[Tensor] = None) -> Tensor:
        """conv_in .. decoders[-1] FiLM; returns the last activation [B, L, base_ch].  The cond pyramid must be current.
        `chain_serial` (device int32, bumped once per chain by the sampler) enables layer chaining (gw_conv_gn3)."""
        sp = self.spec
        if self.generic:
            return self._body_generic(ws, net_a, net_b, step_ptr, film, film_b_stride, film_step_stride)
        self._chain_serial = chain_serial
        ws.chain_prev = None                        # the first block is not a chained producer
        d = sp.depth
        B, Cx, L = net_a.shape
        st = _cabi.stream_ptr()
        Cc0 = sp.cond_in_ch
        n_coef = (self.lib.gw_conv_in_direct_ws_floats(B, Cx, L, sp.base_ch, Cc0)
                  if (ws.stats is None and self.fuse_gn and self.direct_first and self.dtype == "bf16") else 0)
        if n_coef > 0:
            if ws.coef0 is None or ws.coef0.numel() < n_coef:
                ws.coef0 = torch.empty(n_coef, device=self.device, dtype=torch.float32)
            check(self.lib.gw_conv_in_direct(ptr(net_a), ptr(net_b), ptr(step_ptr), B, Cx, L, ptr(self.p["encoders.0.0.weight"]),
                                             ptr(self.p["encoders.0.0.bias"]), sp.base_ch, ptr(self.p["encoders.0.1.weight"]),
                                             ptr(self.p["encoders.0.1.bias"]), Cc0,
                                             ptr(self.p["cond_enc.0.weight"]) if Cc0 > 0 else None,
                                             ptr(self.p["cond_enc.0.bias"]) if Cc0 > 0 else None, ptr(film),
                                             sp.film_offsets()[0], film_b_stride, film_step_stride, ptr(ws.out[0]),
                                             ptr(ws.pooled[0]), ptr(ws.coef0), st), "conv_in_direct")
            self.launches += 2
        elif (ws.stats is None and self.fuse_gn and self.dtype == "bf16" and ws.sync is not None
                and self.lib.gw_conv_in_gn_group(Cx, L, sp.base_ch, Cc0) > 0):
            # inference: conv + GroupNorm + SiLU + cond + FiLM + pool of the first block in one kernel (conv_in_gn.cu)
            check(self.lib.gw_conv_in_gn(ptr(net_a), ptr(net_b), ptr(step_ptr), B, Cx, L, ptr(self.p["encoders.0.0.weight"]),
                                         ptr(self.p["encoders.0.0.bias"]), sp.base_ch, ptr(self.p["encoders.0.1.weight"]),
                                         ptr(self.p["encoders.0.1.bias"]), Cc0,
                                         ptr(self.p["cond_enc.0.weight"]) if Cc0 > 0 else None,
                                         ptr(self.p["cond_enc.0.bias"]) if Cc0 > 0 else None, ptr(film),
                                         sp.film_offsets()[0], film_b_stride, film_step_stride, ptr(ws.out[0]),
                                         ptr(ws.pooled[0]), ptr(ws.sync), st), "conv_in_gn")
            self.launches += 1
        elif ws.stats is None and self.fuse_first_block:
            # inference: the first block never materialises its raw conv output (stats pass + recompute/apply pass)
            Cc = sp.cond_in_ch
            check(self.lib.gw_conv_in_block(ptr(net_a), ptr(net_b), ptr(step_ptr), B, Cx, L, ptr(self.p["encoders.0.0.weight"]),
                                            ptr(self.p["encoders.0.0.bias"]), sp.base_ch, ptr(self.p["encoders.0.1.weight"]),
                                            ptr(self.p["encoders.0.1.bias"]), Cc,
                                            ptr(self.p["cond_enc.0.weight"]) if Cc > 0 else None,
                                            ptr(self.p["cond_enc.0.bias"]) if Cc > 0 else None, ptr(film),
                                            sp.film_offsets()[0], film_b_stride, film_step_stride, ptr(ws.out[0]),
                                            ptr(ws.pooled[0]), self.gw_dtype, ptr(ws.part), st), "conv_in_block")
            self.launches += 2
        else:
            check(self.lib.gw_conv_in(ptr(net_a), ptr(net_b), ptr(step_ptr), B, Cx, L, ptr(self.p["encoders.0.0.weight"]),
                                      ptr(self.p["encoders.0.0.bias"]), sp.base_ch, ptr(ws.raw[0]), self.gw_dtype,
                                      ptr(ws.part), st), "conv_in")
            self.launches += 1
            self._gn(0, ws, (L + 127) // 128, film, film_b_stride, film_step_stride, step_ptr, ws.pooled[0], 0)
        for i in range(1, d):
            self._block(i, ws, ws.pooled[i - 1], None, film, film_b_stride, film_step_stride, step_ptr, ws.pooled[i], i)
        self._block(d, ws, ws.pooled[d - 1], None, film, film_b_stride, film_step_stride, step_ptr, None, d)
        h = ws.out[d]
        ws.head_fused = False
        for i in range(d):
            li = d + 1 + i
            last = i == d - 1
            hf = self._block(li, ws, h, ws.out[d - 1 - i], film, film_b_stride, film_step_stride, step_ptr, None, d - 1 - i,
                             head=last and self.fuse_head)
            h = ws.out[li]
            if last and hf:
                ws.head_fused = True
                h = ws.dots                                   # consumed by head() (gw_final_step with dtype = GW_DOTS)
        return h

    # ------------------------------------------------------------------ shape-generic path (csrc/generic.cu)
    def _block_generic(self, li: int, ws: _Workspace, src0: Optional[Tensor], src1: Optional[Tensor], net_a: Optional[Tensor],
                       net_b: Optional[Tensor], step_ptr: Optional[Tensor], film: Tensor, film_b_stride: int,
                       film_step_stride: int, pooled: Optional[Tensor], lvl: int) -> None:
        """Conv1d(K) + GroupNorm(gcd(8, C)) + SiLU + cond bias + FiLM (+ avg-pool) of layer li (models.py:160-173, 203-208)."""
        sp, lib = self.spec, self.lib
        name, cname = sp.layer_names()[li], sp.cond_names()[li]
        raw, out = ws.raw[li], ws.out[li]
        B, L, Cout = raw.shape
        st = _cabi.stream_ptr()
        Cc = sp.cond_in_ch
        if src0 is None:
            Cx = net_a.shape[1]
            check(lib.gw_gen_conv(None, 0, 0, 0, None, 0, ptr(net_a), ptr(net_b), ptr(step_ptr), Cx, B, L,
                                  ptr(self.p[name + ".0.weight"]), ptr(self.p[name + ".0.bias"]), Cout, sp.kernel, ptr(raw),
                                  self.gw_dtype, st), f"gen_conv[{name}]")
        else:
            C1 = src1.shape[2] if src1 is not None else 0
            check(lib.gw_gen_conv(ptr(src0), src0.shape[2], src0.shape[1], 1 if src1 is not None else 0, ptr(src1), C1, None, None,
                                  None, 0, B, L, ptr(self.p[name + ".0.weight"]), ptr(self.p[name + ".0.bias"]), Cout, sp.kernel,
                                  ptr(raw), self.gw_dtype, st), f"gen_conv[{name}]")
        groups = math.gcd(8, Cout)
        check(lib.gw_gen_gn_stats(ptr(raw), B, L, Cout, groups, self.gw_dtype, ptr(ws.stats[li]), st), f"gen_gn_stats[{name}]")
        check(lib.gw_gen_gn_apply(ptr(raw), ptr(ws.stats[li]), B, L, Cout, groups, ptr(self.p[name + ".1.weight"]),
                                  ptr(self.p[name + ".1.bias"]), ptr(ws.cond[lvl]) if Cc > 0 else None, Cc,
                                  ptr(self.p[cname + ".weight"]) if Cc > 0 else None,
                                  ptr(self.p[cname + ".bias"]) if Cc > 0 else None, ptr(film), sp.film_offsets()[li],
                                  film_b_stride, film_step_stride, ptr(step_ptr), ptr(out), ptr(pooled), self.gw_dtype, st),
              f"gen_gn_apply[{name}]")
        self.launches += 3

    def _body_generic(self, ws: _Workspace, net_a: Tensor, net_b: Optional[Tensor], step_ptr: Optional[Tensor], film: Tensor,
                      film_b_stride: int, film_step_stride: int) -> Tensor:
        d = self.spec.depth
        ws.head_fused = False
        self._block_generic(0, ws, None, None, net_a, net_b, step_ptr, film, film_b_stride, film_step_stride, ws.pooled[0], 0)
        for i in range(1, d):
            self._block_generic(i, ws, ws.pooled[i - 1], None, None, None, step_ptr, film, film_b_stride, film_step_stride,
                                ws.pooled[i], i)
        self._block_generic(d, ws, ws.pooled[d - 1], None, None, None, step_ptr, film, film_b_stride, film_step_stride, None, d)
        for i in range(d):
            li = d + 1 + i
            self._block_generic(li, ws, ws.out[li - 1], ws.out[d - 1 - i], None, None, step_ptr, film, film_b_stride,
                                film_step_stride, None, d - 1 - i)
        return ws.out[2 * d]

    def head(self, h: Tensor, net_a: Tensor, net_b: Optional[Tensor], params: StepParams, coef: Optional[Tensor],
             step_ptr: Optional[Tensor], noise: Optional[Tensor], eps_out: Optional[Tensor], x0_out: Optional[Tensor],
             B: int) -> None:
        _, Cx, L = net_a.shape
        if self.generic:
            check(self.lib.gw_gen_final(ptr(h), self.gw_dtype, ptr(net_a), ptr(net_b), B, Cx, L, self.spec.base_ch, self.spec.kernel,
                                        ptr(self.wf), ptr(self.p["final.bias"]), C.byref(params), ptr(coef), ptr(step_ptr),
                                        ptr(noise), ptr(eps_out), ptr(x0_out), _cabi.stream_ptr()), "gen_final")
            self.launches += 1
            return
        dots = h.dtype == torch.float32 and h.dim() == 3 and h.shape[-1] == 4 and self.dtype == "bf16"
        check(self.lib.gw_final_step(ptr(h), 2 if dots else self.gw_dtype, ptr(net_a), ptr(net_b), B, Cx, L, self.spec.base_ch,
                                     ptr(self.wf), ptr(self.p["final.bias"]), C.byref(params), ptr(coef), ptr(step_ptr),
                                     ptr(noise), ptr(eps_out), ptr(x0_out), _cabi.stream_ptr()), "final_step")
        self.launches += 1

    # ------------------------------------------------------------------ public: plain forward
    @torch.no_grad()
    def forward(self, x: Tensor, t: Tensor, keep_raw: bool = False) -> Tensor:
        """UNet1D.forward (models.py:195-231): x [B, C, L] fp32, t [B] -> eps_hat [B, 1, L] fp32."""
        sp = self.spec
        if x.device.type != "cuda":
            raise RuntimeError("gwb200: CUDA tensors only (no CPU fallback)")
        B, Cx, L = x.shape
        if Cx != sp.in_ch:
            raise ValueError(f"expected {sp.in_ch} input channels, got {Cx}")
        if L < 2 ** sp.depth:
            raise ValueError("sequence too short for the U-Net depth")
        x = x.contiguous().float()
        ws = self.workspace(B, L, keep_raw)
        film = self.film_vectors(t.reshape(-1).expand(B) if t.numel() == 1 else t)
        self.cond_pyramid(ws, x)
        h = self.body(ws, x, None, None, film, sp.film_dim, 0)
        eps = torch.empty(B, 1, L, device=self.device, dtype=torch.float32)
        prm = StepParams(0, 0, 0, 0, 1.0, 0.0, None, 0, 0)
        self.head(h, x, None, prm, None, None, None, eps, None, B)
        return eps


# ======================================================================================================
# sampler
# ======================================================================================================
def build_t_schedule(T: int, steps: int, start_t: Optional[int]) -> List[int]:
    """inference.py:217-228, on the host (pure integers; fp32 linspace + round like the reference)."""
    if start_t is None:
        start_t = T - 1
    start_t = int(max(0, min(start_t, T - 1)))
    steps = int(max(1, min(steps, start_t + 1)))
    ts = torch.linspace(start_t, 0, steps).round().long()
    ts = torch.unique_consecutive(ts)
    out = [int(v) for v in ts]
    if out[0] != start_t:
        out = [start_t] + out
    if out[-1] != 0:
        out = out + [0]
    return out


def cfg_weight(i: int, N: int, mode: str, wmax: float, center: float, width: float) -> float:
    """inference.py:230-244."""
    s = 1.0 if N <= 1 else i / (N - 1)
    mode = mode.lower()
    if mode == "const":
        return float(wmax)
    if mode == "tophat":
        lo, hi = center - width * 0.5, center + width * 0.5
        return float(wmax) if (lo <= s <= hi) else 1.0
    if mode == "gauss":
        sig = max(width, 1e-9)
        return float(wmax) * math.exp(-0.5 * ((s - center) / sig) ** 2)
    raise ValueError(f"unknown cfg-mode: {mode}")


class SamplerPlan:
    """One DDIM/DDPM chain configuration bound to (B, L): device tables + a captured CUDA graph."""

    def __init__(self, eng: UNetEngine, alpha_bar: Tensor, B: int, L: int, *, T: int, steps: int, eta: float,
                 start_t: Optional[int], dc_weight: float, eps_scale: float, pred_type: str, cfg_scale: float,
                 cfg_mode: str, cfg_center: float, cfg_width: float, cfg_u_only_thresh: float, seed: int = 0,
                 sample0: int = 0):
        self.eng, self.B, self.L = eng, B, L
        sp = eng.spec
        dev = eng.device
        sched = build_t_schedule(T, steps, start_t)
        N = len(sched)
        self.sched, self.N = sched, N
        ab = alpha_bar.detach().float().cpu().clamp(1e-12, 1.0)           # inference.py:399
        coef = torch.zeros(N, 16, dtype=torch.float32)
        uses = []
        draw = 1                                                        # draw 0 initialises x_T
        for i, t_now in enumerate(sched):
            ab_t = ab[t_now]
            ab_prev = ab[sched[i + 1]] if i + 1 < N else torch.tensor(1.0)
            w = cfg_weight(i, N, cfg_mode, cfg_scale, cfg_center, cfg_width)
            if w <= cfg_u_only_thresh:
                use = 1
            elif abs(w - 1.0) <= 1e-6:
                use = 0
            else:
                use = 2
            uses.append(use)
            sig = eta * torch.sqrt((1 - ab_prev) / (1 - ab_t) * (1 - ab_t / ab_prev))       # inference.py:481
            dirc = torch.sqrt(torch.clamp(1 - ab_prev - sig ** 2, min=0.0))                 # inference.py:482
            last = 1.0 if t_now == 0 else 0.0
            sigv = float(sig) if (not last and float(sig) > 0 and math.isfinite(float(sig))) else 0.0
            coef[i, 0] = torch.sqrt(1 - ab_t)
            coef[i, 1] = torch.sqrt(ab_t)
            coef[i, 2] = torch.sqrt(ab_prev)
            coef[i, 3] = dirc if math.isfinite(float(dirc)) else 0.0
            coef[i, 4] = sigv
            coef[i, 5] = w
            coef[i, 6] = float(use)
            coef[i, 7] = last
            coef[i, 8] = float(draw if sigv > 0 else 0)
            coef[i, 9] = torch.sqrt(torch.clamp(1 - ab_t, min=1e-12))
            if sigv > 0:
                draw += 1
        self.n_draws = draw
        self.ab_start = float(ab[sched[0]])
        self.cfg_both = any(u != 0 for u in uses)       # an unconditional batch is needed at some step
        self.Bn = B * (2 if self.cfg_both else 1)
        self.coef = coef.to(dev)
        self.sched_dev = torch.tensor(sched, dtype=torch.int64, device=dev)
        self.film = eng.film_vectors(self.sched_dev)
        self._gen = eng.generation
        self._ptr_gen = eng.ptr_generation
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        Cx = sp.in_ch
        self.net = [torch.zeros(self.Bn, Cx, L, device=dev, dtype=torch.float32) for _ in range(2)]
        self.ws = eng.workspace(self.Bn, L)
        self.y_dc = torch.zeros(B, L, device=dev, dtype=torch.float32) if dc_weight > 0 else None
        self.adv = torch.zeros(1, dtype=torch.int32, device=dev)      # CTA counter of the self-advancing head kernel
        self.serial = torch.zeros(1, dtype=torch.int32, device=dev)   # chain serial: (serial, step) tags the layer-chaining flags
        if N >= 4096:
            raise ValueError("SamplerPlan: at most 4095 reverse steps per chain")
        # Philox key {seed, sample0} lives in device memory: the kernel arguments of a captured graph are frozen, the key of
        # a cached plan is not (chunked sweeps, rank shards, repeated seed=None calls)
        self.rng = torch.zeros(2, dtype=torch.int64, device=dev)
        self._rng_host = torch.zeros(2, dtype=torch.int64).pin_memory()
        self._rng_ev: Optional[torch.cuda.Event] = None
        self.params = StepParams(1, 1 if self.cfg_both else 0, 1 if sp.use_selfcond else 0, 1 if pred_type != "eps" else 0,
                                 float(eps_scale), float(dc_weight), ptr(self.y_dc), int(seed) & (2 ** 64 - 1), int(sample0),
                                 ptr(self.adv), ptr(self.rng))
        self.set_rng(seed, sample0)
        self.noise: Optional[Tensor] = None
        self.trace_eps: Optional[Tensor] = None
        self.trace_x0: Optional[Tensor] = None
        self._graph = None
        self.graph_capture_ms: Optional[float] = None
        self._graph_steps = 0

    def set_rng(self, seed: int, sample0: int) -> None:
        """Philox key of the stochastic steps (stream-ordered copy into the device key the kernels read)."""
        seed = int(seed) & (2 ** 64 - 1)
        if self._rng_ev is not None:
            self._rng_ev.synchronize()                 # the previous copy out of the pinned pair has completed
        self._rng_host[0] = seed - (1 << 64) if seed >= (1 << 63) else seed
        self._rng_host[1] = int(sample0)
        self.rng.copy_(self._rng_host, non_blocking=True)
        self._rng_ev = torch.cuda.Event()
        self._rng_ev.record()
        self.params.seed, self.params.sample0 = seed, int(sample0)

    def sync_weights(self) -> None:
        """Recompute the FiLM table (in place: a captured graph stays valid) when the engine's weights were refreshed."""
        if self._gen != self.eng.generation:
            self.eng.film_vectors(self.sched_dev, out=self.film)
            self._gen = self.eng.generation
        if self._ptr_gen != self.eng.ptr_generation:   # the captured kernel arguments point at the parameters' OLD storage
            self._graph = None
            self._ptr_gen = self.eng.ptr_generation

    # one reverse step = 8 kernels (bf16 fused path)
    def enqueue_step(self) -> None:
        eng = self.eng
        h = eng.body(self.ws, self.net[0], self.net[1], self.step, self.film, 0, eng.spec.film_dim, chain_serial=self.serial)
        eng.head(h, self.net[0], self.net[1], self.params, self.coef, self.step, self.noise, self.trace_eps, self.trace_x0, self.B)
        if not self.ws.head_fused:                     # the head kernel on the fused dots advances the counter itself
            check(eng.lib.gw_step_advance(ptr(self.step), -1, _cabi.stream_ptr()), "step_advance")
            eng.launches += 1

    def load_inputs(self, x_init: Tensor, cond_on: Tensor, cond_off: Optional[Tensor], y_dc: Optional[Tensor]) -> None:
        """x_init [B,1,L]; cond_on/off [B,Cc,L] (already cond-scaled / zeroed as inference.py:434-446)."""
        sp, B = self.eng.spec, self.B
        self.sync_weights()
        for net in self.net:
            net.zero_()
            net[:B, 0:1] = x_init
            if sp.cond_in_ch > 0:
                net[:B, 1:1 + sp.cond_in_ch] = cond_on
            if self.cfg_both:
                net[B:, 0:1] = x_init
                if sp.cond_in_ch > 0 and cond_off is not None:
                    net[B:, 1:1 + sp.cond_in_ch] = cond_off
        if self.y_dc is not None:
            self.y_dc.copy_(y_dc.reshape(B, self.L))
        self.step.zero_()
        self.serial.add_(1)                            # a new chain: its (serial, step) tags differ from every earlier chain's
        self.eng.cond_pyramid(self.ws, self.net[0])

    def capture(self, steps_per_graph: int) -> None:
        """Capture `steps_per_graph` reverse steps in one CUDA graph (N for the whole chain)."""
        if self._graph is not None and self._graph_steps == steps_per_graph:
            return
        # warm-up outside capture (lazy packing, cudaFuncSetAttribute) on a side stream, as torch requires
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        saved = [n.clone() for n in self.net]
        with torch.cuda.stream(s):
            self.enqueue_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for n, sv in zip(self.net, saved):
            n.copy_(sv)
        self.step.zero_()
        self.serial.add_(1)                            # the warm-up step used (serial, step 0): its chain flags must not match again
        import time
        t0 = time.perf_counter()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(steps_per_graph):
                self.enqueue_step()
        torch.cuda.synchronize()
        self.graph_capture_ms = (time.perf_counter() - t0) * 1e3     # stream capture + cudaGraphInstantiate (host wall clock)
        self._graph, self._graph_steps = g, steps_per_graph
        self.step.zero_()

    def run(self, use_graph: bool = True, steps_per_graph: Optional[int] = None) -> Tensor:
        """Run the N-step chain from the loaded inputs; returns x_0 estimate [B,1,L] (a view of the ping-pong buffer)."""
        N = self.N
        if use_graph:
            spg = steps_per_graph or N
            if N % spg != 0:
                spg = 1
            self.capture(spg)
            for _ in range(N // spg):
                self._graph.replay()
        else:
            for _ in range(N):
                self.enqueue_step()
        return self.net[N & 1][: self.B, 0:1]
