"""Drop-in mirror of the reference training hot path (src/snr_denoising/train.py:40-172, 320-456) on the sm_100a kernels.

Three layers, from the reference-facing one down:
  * the reference helper names (`_predict_x0_norm`, `_element_loss`, `update_ema`, `make_warmup_cosine_scheduler`,
    `_match_batch`, `_sample_timesteps_stratified`) and `train_diffusion(args)`;
  * `unet_autograd_forward`: `UNet1D.forward` under autograd -- `loss.backward()` in a reference-style loop runs the
    hand-written backward kernels and fills `param.grad`;
  * `FusedTrainStep`: the whole per-batch body (q_sample, CFG dropout, optional self-conditioning forward, forward,
    masked Huber/MSE loss, backward, global-norm clip, AdamW, EMA) as one kernel sequence over flat parameter / gradient /
    optimiser buffers, CUDA-graph captured, with one NCCL all-reduce of the flat gradient bucket when world_size > 1.
Everything numerical is in libgwb200.so; torch supplies memory, streams, graphs and `torch.distributed`.
Dataset / HDF5 / logging parts of the reference file (train.py:17-27, 93-130, 202-217, 467-630) are outside the hot path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import _cabi
from ._cabi import ConvTcShape, StepParams, check, ptr
from .engine import ModelSpec, ParamLayout, UNetEngine, _Workspace

Tensor = torch.Tensor

__all__ = ["BackwardEngine", "FusedTrainStep", "unet_autograd_forward", "_predict_x0_norm", "_element_loss", "update_ema",
           "make_warmup_cosine_scheduler", "warmup_cosine_lambda", "_match_batch", "_sample_timesteps_stratified",
           "train_diffusion"]


# ======================================================================================================
# backward engine
# ======================================================================================================
class _GradWorkspace:
    """Activation-gradient buffers for one (B, L): sized once, reused by every layer (see BackwardEngine.backward)."""

    def __init__(self, spec: ModelSpec, ws: _Workspace, tdtype, device, lib, simt: bool):
        B = ws.B
        d = spec.depth
        lc = spec.layer_channels
        per = [ws.lay_len[i] * lc[i] for i in range(2 * d + 1)]
        mx = max(per)
        e = lambda n: torch.empty(B * n, device=device, dtype=tdtype)
        self.d_raw = e(mx)
        self.d_h = [e(mx), e(mx)]
        self.d_skip = [e(per[i]) for i in range(d)]
        self.d_pool = [e(mx), e(mx)]
        cat_max = max(ws.lay_len[d + 1 + i] * (lc[d + i] + spec.chs[d - 1 - i]) for i in range(d))
        self.d_cat = e(cat_max) if simt else None
        n_scr = max(lib.gw_gn_bwd_scratch_elems(B, ws.lay_len[i], lc[i], spec.cond_in_ch) for i in range(2 * d + 1))
        n_scr = max(n_scr, B * ((ws.L + 511) // 512) * ((lc[-1] + 1) * 3 + 1), B * spec.base_ch * (1 + (spec.film_dim + 95) // 96),
                    B * ((ws.L + 511) // 512) * lc[0] * spec.in_ch * 3)
        wmax = max(lc[i] * ((lc[i - 1] + (spec.chs[2 * d - i] if i > d else 0)) * 3) for i in range(1, 2 * d + 1))
        self.wg_elems = max(16 * wmax, 8 << 20)                           # split-K partials of the SIMT / tcgen05 wgrad
        for i in range(1, 2 * d + 1):
            cin_parts = [(0, lc[i - 1])] if i <= d else [(1, lc[i - 1]), (0, spec.chs[2 * d - i])]
            for mode, cx in cin_parts:
                if mode == 1 and ws.lay_len[i] % 2:
                    continue
                self.wg_elems = max(self.wg_elems, lib.gw_wgrad_tc_scratch_elems(mode, B, ws.lay_len[i], lc[i], cx))
        n_scr = max(n_scr, lib.gw_gen_gn_bwd_scratch_floats(B, max(lc)))   # generic path: per-(sample, channel) sums
        self.scratch = torch.empty(max(n_scr, self.wg_elems), device=device, dtype=torch.float32)
        self.dfilm = torch.zeros(B, spec.film_dim, device=device, dtype=torch.float32)
        self.aux = torch.empty(B, spec.time_dim + 3 * spec.base_ch, device=device, dtype=torch.float32)
        self.film = torch.empty(B, spec.film_dim, device=device, dtype=torch.float32)
        # exchange buffer of the one-pass GroupNorm backward (gn_bwd_fused.cu): zeroed once, epochs advance per launch
        self.sync = torch.zeros(lib.gw_gn_bwd_sync_bytes(B), device=device, dtype=torch.uint8)
        # deferred small reductions (BackwardEngine.defer_small): the parameter-gradient kernel of a GroupNorm block and the
        # fold / scatter passes of a weight gradient run on a second stream, so what they read must outlive the next launches on
        # the main stream: one scratch buffer per layer / per wgrad call instead of the shared one (a few hundred MB of 180 GB)
        self.gn_scr = [torch.empty(lib.gw_gn_bwd_scratch_elems(B, ws.lay_len[i], lc[i], spec.cond_in_ch), device=device,
                                   dtype=torch.float32) if lc[i] % 64 == 0 else None for i in range(2 * d + 1)]
        self.wg_scr: Dict[tuple, Tensor] = {}
        self.film_scr = torch.empty(B * spec.base_ch * (2 + (spec.film_dim + 95) // 96), device=device, dtype=torch.float32)
        self._dev = device

    def wgrad_scratch(self, key: tuple, n: int) -> Tensor:
        buf = self.wg_scr.get(key)
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, device=self._dev, dtype=torch.float32)
            self.wg_scr[key] = buf
        return buf


class BackwardEngine:
    """Forward-with-saved-activations and backward of UNet1D on top of `UNetEngine`.

    Gradients are ACCUMULATED into `flat_grad` (fp32, `ParamLayout` order; the caller zeroes it once per step)."""

    def __init__(self, eng: UNetEngine, layout: ParamLayout):
        self.eng = eng
        self.lib = eng.lib
        self.layout = layout
        self._gws: Dict[tuple, _GradWorkspace] = {}
        self._wt: Dict[int, Tensor] = {}
        self._dg_packed: Dict[tuple, Tensor] = {}
        # conv backward kernels: tcgen05 GEMMs when the engine runs the tcgen05 forward, CUDA-core fp32 otherwise
        self.dgrad_impl = "tc" if eng.conv_impl == "tc" else "simt"
        self.wgrad_impl = "tc" if eng.conv_impl == "tc" else "simt"
        self.wgrad_variant = 0
        self.fuse_head = False                       # see backward(): measured a wash on B200, the materialised d_h stays the default
        self._prepped = False                        # True while the caller has already run prepare_dgrad() for this step
        # bf16: one-pass GroupNorm backward (gn_bwd_fused.cu).  Reads every operand once (3.5 instead of 5.5 tensor passes) but
        # measured slower on B200 (1.90 vs 1.13 ms per step at B=256, L=4096): a slice stays in shared memory for load +
        # sums + exchange + apply + store (~8 us), and 228 KB per SM cannot cover that latency at HBM rate.  Parity-tested option.
        self.fuse_gn_bwd = False
        # bf16 / tcgen05: the small latency-bound reductions of the backward pass -- the per-block parameter-gradient kernel
        # (gn_bwd_param_kernel, 2 - 8 CTAs, ~15 us), the split-K fold + scatter of every weight gradient (~12 us per call) and the
        # time-MLP backward (4 launches, ~80 us) -- run on a second stream (a parallel branch of the step's CUDA graph) next to
        # the big kernels instead of between them; joined before the gradient norm.  GWB200_DEFER=0 turns it off.
        self.defer_small = os.environ.get("GWB200_DEFER", "1") != "0"
        self._side2: Optional[torch.cuda.Stream] = None
        self._deferring = False

    def grad_workspace(self, ws: _Workspace) -> _GradWorkspace:
        key = (ws.B, ws.L)
        g = self._gws.get(key)
        if g is None:
            g = _GradWorkspace(self.eng.spec, ws, self.eng.tdtype, self.eng.device, self.lib, simt=True)
            self._gws[key] = g
        return g

    # ------------------------------------------------------------------ forward (training mode)
    def forward(self, net: Tensor, t: Tensor, eps_out: Optional[Tensor] = None) -> Tensor:
        """UNet1D.forward keeping raw conv outputs, GroupNorm statistics and the time-MLP activations for `backward`."""
        eng, sp = self.eng, self.eng.spec
        B, Cx, L = net.shape
        ws = eng.workspace(B, L, True)
        g = self.grad_workspace(ws)
        eng.film_vectors(t, out=g.film, aux=g.aux)
        eng.cond_pyramid(ws, net)
        h = eng.body(ws, net, None, None, g.film, sp.film_dim, 0)
        if eps_out is None:
            eps_out = torch.empty(B, 1, L, device=eng.device, dtype=torch.float32)
        eng.head(h, net, None, StepParams(0, 0, 0, 0, 1.0, 0.0, None, 0, 0), None, None, None, eps_out, None, B)
        return eps_out

    # ------------------------------------------------------------------ backward
    # ------------------------------------------------------------------ dgrad weights (depend on the parameters only)
    def _layer_io(self, li: int, ws: _Workspace):
        d = self.eng.spec.depth
        if li <= d:
            return ws.pooled[li - 1], None
        return ws.out[li - 1], ws.out[2 * d - li]

    def _dgrad_tc_ok(self, li: int, ws: _Workspace) -> bool:
        src0, src1 = self._layer_io(li, ws)
        _, L, Cout = ws.raw[li].shape
        C0, L0 = src0.shape[2], src0.shape[1]
        C1 = src1.shape[2] if src1 is not None else 0
        return (self.eng.dtype == "bf16" and (src1 is None or (L % 2 == 0 and L0 * 2 == L)) and Cout <= 256 and C0 <= 256
                and C1 <= 256)

    def _dgrad_parts(self, li: int, ws: _Workspace):
        """(conv shape, first row of the dgrad weight matrix) per dgrad GEMM of layer li: plain conv for pooled / skip inputs,
        pair-sum mode through the nearest upsample."""
        src0, src1 = self._layer_io(li, ws)
        B, L, Cout = ws.raw[li].shape
        C0 = src0.shape[2]

        def plain(cout_):                     # 64 output channels: evaluate in pair space to fill a 128-column tile
            return ConvTcShape(1, 3 if (cout_ == 64 and L % 2 == 0) else 0, B, L, Cout, L, 0, cout_)
        if src1 is None:
            return [(plain(C0), 0)]
        return [(ConvTcShape(1, 2, B, L // 2, Cout, L, 0, C0), 0), (plain(src1.shape[2]), C0)]

    def _prep_dgrad_layer(self, li: int, ws: _Workspace) -> None:
        """w'[ci][co][k] = w[co][ci][2-k] (+ the bf16 GEMM packing for the tcgen05 dgrad) of conv li."""
        eng, lib = self.eng, self.lib
        st = _cabi.stream_ptr()
        name = eng.spec.layer_names()[li]
        src0, src1 = self._layer_io(li, ws)
        _, L, Cout = ws.raw[li].shape
        Cin = src0.shape[2] + (src1.shape[2] if src1 is not None else 0)
        w = eng.p[name + ".0.weight"]
        wt = self._wt.get(li)
        if wt is None:
            wt = torch.empty(w.numel(), device=eng.device, dtype=torch.float32)
            self._wt[li] = wt
        if eng.generic:
            check(lib.gw_gen_weight_dgrad(ptr(w), Cout, Cin, eng.spec.kernel, ptr(wt), st), "gen_weight_dgrad")
            eng.launches += 1
            return
        check(lib.gw_weight_dgrad(ptr(w), Cout, Cin, ptr(wt), st), "weight_dgrad")
        eng.launches += 1
        if self.dgrad_impl == "tc" and self._dgrad_tc_ok(li, ws):
            for pi, (shp, row0) in enumerate(self._dgrad_parts(li, ws)):
                key = (li, pi, L)
                packed = self._dg_packed.get(key)
                if packed is None:
                    n = lib.gw_conv_tc_packed_elems(C.byref(shp))
                    if n <= 0:
                        raise RuntimeError("gw_conv_tc_packed_elems(dgrad): " + lib.gw_last_error().decode())
                    packed = torch.empty(n, device=eng.device, dtype=torch.bfloat16)
                    self._dg_packed[key] = packed
                check(lib.gw_conv_tc_pack(C.byref(shp), wt.data_ptr() + 4 * row0 * Cout * 3, ptr(packed), st), "conv_tc_pack(dgrad)")
                eng.launches += 1

    def prepare_dgrad(self, ws: _Workspace) -> None:
        """All layers' dgrad weights on the CURRENT stream.  They depend on the parameters only, so a caller may run this on a
        side stream next to the forward pass and set `_prepped` around `backward` (FusedTrainStep._enqueue)."""
        for li in range(1, 2 * self.eng.spec.depth + 1):
            self._prep_dgrad_layer(li, ws)

    def _conv_bwd(self, li: int, ws: _Workspace, g: _GradWorkspace, grads: Dict[str, Tensor], d_in0: Tensor,
                  d_in1: Optional[Tensor]) -> None:
        """wgrad + dgrad of conv `li` (>= 1) given g.d_raw.  d_in0 receives the gradient wrt src0 (the pooled tensor, or h
        before the nearest upsample), d_in1 the gradient wrt the skip (decoders)."""
        eng, sp, lib = self.eng, self.eng.spec, self.lib
        d = sp.depth
        name = sp.layer_names()[li]
        B, L, Cout = ws.raw[li].shape
        st = _cabi.stream_ptr()
        if li <= d:
            src0, src1, up = ws.pooled[li - 1], None, 0
        else:
            src0, src1, up = ws.out[li - 1], ws.out[2 * d - li], 1
        C0, L0 = src0.shape[2], src0.shape[1]
        C1 = src1.shape[2] if src1 is not None else 0
        tc_ok = eng.dtype == "bf16" and (src1 is None or (L % 2 == 0 and L0 * 2 == L)) and Cout <= 256 and C0 <= 256 and C1 <= 256
        dW = grads[name + ".0.weight"]
        if eng.generic:
            K = sp.kernel
            check(lib.gw_gen_wgrad(ptr(src0), C0, L0, up, ptr(src1), C1, None, 0, ptr(g.d_raw), B, L, Cout, K, eng.gw_dtype,
                                   ptr(dW), st), f"gen_wgrad[{name}]")
            if not self._prepped:
                self._prep_dgrad_layer(li, ws)
            dst = d_in0 if src1 is None else g.d_cat
            check(lib.gw_gen_conv(ptr(g.d_raw), Cout, L, 0, None, 0, None, None, None, 0, B, L, ptr(self._wt[li]), None, C0 + C1, K,
                                  ptr(dst), eng.gw_dtype, st), f"gen_dgrad[{name}]")
            eng.launches += 2
            if src1 is not None:
                check(lib.gw_gen_split_cat(ptr(g.d_cat), B, L, C0, L0, C1, ptr(d_in0), ptr(d_in1), eng.gw_dtype, st), "gen_split_cat")
                eng.launches += 1
            return
        if self.wgrad_impl == "tc" and tc_ok:
            defer = self._deferring and (self.wgrad_variant & 2) == 0
            parts = [(0, src0, C0, C0, 0)] if src1 is None else [(1, src0, C0, C0 + C1, 0), (0, src1, C1, C0 + C1, C0)]
            for pi, (mode, src, cx, ctot, off) in enumerate(parts):
                if not defer:
                    check(lib.gw_wgrad_tc(mode, ptr(g.d_raw), ptr(src), B, L, Cout, cx, ctot, off, ptr(g.scratch), g.scratch.numel(),
                                          ptr(dW), self.wgrad_variant, st), f"wgrad_tc[{name}.{pi}]")
                else:
                    # GEMM on the main stream into this call's own partial buffer; fold + scatter-accumulate on the side stream
                    scr = g.wgrad_scratch((li, pi), lib.gw_wgrad_tc_scratch_elems(mode, B, L, Cout, cx))
                    check(lib.gw_wgrad_tc(mode, ptr(g.d_raw), ptr(src), B, L, Cout, cx, ctot, off, ptr(scr), scr.numel(), ptr(dW),
                                          self.wgrad_variant | 4, st), f"wgrad_tc[{name}.{pi}]")
                    self._side2.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(self._side2):
                        check(lib.gw_wgrad_tc_finish(mode, B, L, Cout, cx, ctot, off, ptr(scr), ptr(dW), _cabi.stream_ptr()),
                              f"wgrad_tc_finish[{name}.{pi}]")
                eng.launches += 2
        else:
            check(lib.gw_wgrad3_simt(ptr(src0), C0, L0, up, ptr(src1), C1, ptr(g.d_raw), B, L, Cout, eng.gw_dtype,
                                     ptr(g.scratch), g.wg_elems, ptr(dW), st), f"wgrad3_simt[{name}]")
            eng.launches += 2
        # dgrad = the same conv with flipped / transposed weights (conv_transpose of a stride-1 'same' conv)
        if not self._prepped:
            self._prep_dgrad_layer(li, ws)
        wt = self._wt[li]
        if self.dgrad_impl == "tc" and tc_ok:
            dsts = [d_in0] if src1 is None else [d_in0, d_in1]
            for pi, (shp, row0) in enumerate(self._dgrad_parts(li, ws)):
                packed = self._dg_packed[(li, pi, L)]
                variant = eng.tc_variant if (shp.Cout >= 128 or shp.pair == 3) else 0
                check(lib.gw_conv_tc(C.byref(shp), ptr(g.d_raw), None, ptr(packed), None, ptr(dsts[pi]), None, variant, st),
                      f"dgrad_tc[{name}.{pi}]")
                eng.launches += 1
            return
        dst = d_in0 if src1 is None else g.d_cat
        check(lib.gw_conv3_simt(ptr(g.d_raw), Cout, L, 0, None, 0, B, L, ptr(wt), None, C0 + C1, ptr(dst), eng.gw_dtype, None,
                                st), f"dgrad_simt[{name}]")
        eng.launches += 2
        if src1 is not None:
            check(lib.gw_split_cat_grad(ptr(g.d_cat), B, L, C0, L0, C1, ptr(d_in0), ptr(d_in1), eng.gw_dtype, st), "split_cat_grad")
            eng.launches += 1

    def backward(self, net: Tensor, d_eps: Tensor, flat_grad: Tensor) -> None:
        """Backward of the forward last run by `forward(net, t)`.  d_eps [B, L] fp32 = dLoss/d eps_hat."""
        eng, sp, lib = self.eng, self.eng.spec, self.lib
        d = sp.depth
        lc = sp.layer_channels
        B, Cx, L = net.shape
        ws = eng.workspace(B, L, True)
        g = self.grad_workspace(ws)
        st = _cabi.stream_ptr()
        grads = self.layout.views(flat_grad)
        names, cnames, foffs = sp.layer_names(), sp.cond_names(), sp.film_offsets()
        Cc = sp.cond_in_ch
        nl = 2 * d + 1
        self._deferring = self.defer_small and eng.dtype == "bf16" and not eng.generic and not self.fuse_gn_bwd
        if self._deferring and self._side2 is None:
            self._side2 = torch.cuda.Stream()

        def gn_bwd(li: int, do_a: Optional[Tensor], do_pool: Optional[Tensor], do_eps: Optional[Tensor] = None) -> None:
            n = names[li]
            lvl = li if li <= d else 2 * d - li
            _, Ll, Cl = ws.raw[li].shape
            if eng.generic:
                check(lib.gw_gen_gn_bwd(ptr(ws.raw[li]), ptr(ws.stats[li]), B, Ll, Cl, math.gcd(8, Cl), ptr(eng.p[n + ".1.weight"]),
                                        ptr(eng.p[n + ".1.bias"]), ptr(ws.cond[lvl]) if Cc > 0 else None, Cc,
                                        ptr(eng.p[cnames[li] + ".weight"]) if Cc > 0 else None,
                                        ptr(eng.p[cnames[li] + ".bias"]) if Cc > 0 else None, ptr(g.film), foffs[li], sp.film_dim,
                                        ptr(do_a), ptr(do_pool), eng.gw_dtype, ptr(g.scratch), ptr(g.dfilm), sp.film_dim,
                                        ptr(g.d_raw), ptr(grads[n + ".1.weight"]), ptr(grads[n + ".1.bias"]),
                                        ptr(grads[cnames[li] + ".weight"]) if Cc > 0 else None,
                                        ptr(grads[cnames[li] + ".bias"]) if Cc > 0 else None, ptr(grads[n + ".0.bias"]), st),
                      f"gen_gn_bwd[{n}]")
                eng.launches += 3
                return
            if self._deferring and do_eps is None and g.gn_scr[li] is not None:
                def call(phase):
                    check(lib.gw_gn_bwd_phase(ptr(ws.raw[li]), ptr(ws.stats[li]), B, Ll, Cl, ptr(eng.p[n + ".1.weight"]),
                                              ptr(eng.p[n + ".1.bias"]), ptr(ws.cond[lvl]) if Cc > 0 else None, Cc,
                                              ptr(eng.p[cnames[li] + ".weight"]) if Cc > 0 else None,
                                              ptr(eng.p[cnames[li] + ".bias"]) if Cc > 0 else None, ptr(g.film), foffs[li],
                                              sp.film_dim, ptr(do_a), ptr(do_pool), None, None, eng.gw_dtype, ptr(g.gn_scr[li]),
                                              ptr(g.dfilm), sp.film_dim, ptr(g.d_raw), ptr(grads[n + ".1.weight"]),
                                              ptr(grads[n + ".1.bias"]), ptr(grads[cnames[li] + ".weight"]) if Cc > 0 else None,
                                              ptr(grads[cnames[li] + ".bias"]) if Cc > 0 else None, ptr(grads[n + ".0.bias"]),
                                              phase, _cabi.stream_ptr()), f"gn_bwd[{n}].{phase}")
                call(1)                                       # statistics, fold, apply: d_raw and the FiLM gradient rows
                self._side2.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side2):
                    call(2)                                   # parameter / conv-bias gradients of this block
                eng.launches += 5
                return
            check(lib.gw_gn_bwd2(ptr(ws.raw[li]), ptr(ws.stats[li]), B, Ll, Cl, ptr(eng.p[n + ".1.weight"]),
                                ptr(eng.p[n + ".1.bias"]), ptr(ws.cond[lvl]) if Cc > 0 else None, Cc,
                                ptr(eng.p[cnames[li] + ".weight"]) if Cc > 0 else None,
                                ptr(eng.p[cnames[li] + ".bias"]) if Cc > 0 else None, ptr(g.film), foffs[li], sp.film_dim,
                                ptr(do_a), ptr(do_pool), ptr(do_eps), ptr(eng.wf) if do_eps is not None else None, eng.gw_dtype,
                                ptr(g.scratch), ptr(g.dfilm), sp.film_dim, ptr(g.d_raw),
                                ptr(grads[n + ".1.weight"]), ptr(grads[n + ".1.bias"]),
                                ptr(grads[cnames[li] + ".weight"]) if Cc > 0 else None,
                                ptr(grads[cnames[li] + ".bias"]) if Cc > 0 else None, ptr(grads[n + ".0.bias"]),
                                ptr(g.sync) if self.fuse_gn_bwd else None, st),
                  f"gn_bwd[{n}]")
            eng.launches += 5

        # The gradient wrt the last block's output is a 3-tap outer product of d_eps and final.weight.  With the specialised
        # streaming GroupNorm-backward kernels (HEAD variants, stream_gn.cu) gw_gn_bwd forms it on the fly from d_eps in shared
        # memory, so the [B, L, 64] d_h tensor is neither written by gw_final_bwd nor read back twice (bf16, `fuse_head`).
        # Measured at B=256, L=4096: gw_final_bwd 130 -> 75 us, but the last block's two GroupNorm-backward passes 162 -> 226 us
        # (they are issue-bound: 12 extra instructions per (row, 4 channels) cost more than the 268 MB of reads they save), so
        # the step time is unchanged (3.68 vs 3.69 ms) and the option stays off.
        fuse_head = self.fuse_head and eng.dtype == "bf16" and not eng.generic
        if eng.generic:
            check(lib.gw_gen_final_bwd(ptr(d_eps), ptr(ws.out[nl - 1]), eng.gw_dtype, ptr(net), B, Cx, L, lc[-1], sp.kernel,
                                       ptr(eng.wf), ptr(g.d_h[0]), ptr(grads["final.weight"]), ptr(grads["final.bias"]), st),
                  "gen_final_bwd")
            eng.launches += 2
        else:
            check(lib.gw_final_bwd(ptr(d_eps), ptr(ws.out[nl - 1]), eng.gw_dtype, ptr(net), B, Cx, L, lc[-1], ptr(eng.wf),
                                   None if fuse_head else ptr(g.d_h[0]), ptr(g.scratch), ptr(grads["final.weight"]),
                                   ptr(grads["final.bias"]), st), "final_bwd")
            eng.launches += 3
        cur = 0
        for li in range(nl - 1, d, -1):                       # decoders, last first
            if li == nl - 1 and fuse_head:
                gn_bwd(li, None, None, d_eps)
            else:
                gn_bwd(li, g.d_h[cur], None)
            self._conv_bwd(li, ws, g, grads, g.d_h[cur ^ 1], g.d_skip[2 * d - li])
            cur ^= 1
        gn_bwd(d, g.d_h[cur], None)                           # mid
        pc = 0
        self._conv_bwd(d, ws, g, grads, g.d_pool[pc], None)
        for li in range(d - 1, 0, -1):                        # encoders d-1 .. 1: skip gradient + avg_pool gradient
            gn_bwd(li, g.d_skip[li], g.d_pool[pc])
            self._conv_bwd(li, ws, g, grads, g.d_pool[pc ^ 1], None)
            pc ^= 1
        gn_bwd(0, g.d_skip[0], g.d_pool[pc])
        lo = self.layout
        cur_s = torch.cuda.current_stream()

        def film_bwd(scr):
            check(lib.gw_film_bwd(ptr(g.dfilm), ptr(g.aux), ptr(eng.film_w2), B, sp.time_dim, sp.base_ch, sp.film_dim, ptr(scr),
                                  ptr(grads["time_mlp.1.weight"]), ptr(grads["time_mlp.1.bias"]),
                                  flat_grad.data_ptr() + 4 * lo.w2[0], flat_grad.data_ptr() + 4 * lo.b2[0], _cabi.stream_ptr()),
                  "film_bwd")
        if self._deferring:
            # dfilm is complete once the first block's GroupNorm backward (phase 1) has run: the time-MLP backward runs on the
            # side stream next to the first conv's weight gradient
            self._side2.wait_stream(cur_s)
            with torch.cuda.stream(self._side2):
                film_bwd(g.film_scr)
        if eng.generic:
            check(lib.gw_gen_wgrad(None, 0, 0, 0, None, 0, ptr(net), Cx, ptr(g.d_raw), B, L, lc[0], sp.kernel, eng.gw_dtype,
                                   ptr(grads["encoders.0.0.weight"]), st), "gen_wgrad[encoders.0]")
            eng.launches += 1
        else:
            check(lib.gw_wgrad_in(ptr(net), B, Cx, L, ptr(g.d_raw), lc[0], eng.gw_dtype, ptr(g.scratch), g.scratch.numel(),
                                  ptr(grads["encoders.0.0.weight"]), st), "wgrad_in")
            eng.launches += 2
        if self._deferring:
            cur_s.wait_stream(self._side2)                    # join: every deferred reduction precedes the gradient norm
        else:
            film_bwd(g.scratch)
        eng.launches += 3


# ======================================================================================================
# autograd bridge: reference-style loops (`loss.backward()`, torch optimisers) on the CUDA kernels
# ======================================================================================================
class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, cd, x, t, *params):
        bwd = model._backward_engine(cd)
        net = x.detach().contiguous().float()
        eps = bwd.forward(net, t.reshape(-1).expand(net.shape[0]) if t.numel() == 1 else t)
        ctx.bwd, ctx.net, ctx.n_params = bwd, net, len(params)
        # the saved activations live in the engine workspace shared per (B, L): stamp it, so a second grad-enabled forward of
        # the same shape before this one's backward is an error instead of silently wrong gradients
        ws = bwd.eng.workspace(net.shape[0], net.shape[2], True)
        ws.fwd_gen = getattr(ws, "fwd_gen", 0) + 1
        ctx.ws, ctx.gen = ws, ws.fwd_gen
        ctx.names = [k for k, _ in model.named_parameters()]
        return eps

    @staticmethod
    def backward(ctx, d_eps):
        bwd = ctx.bwd
        if ctx.ws.fwd_gen != ctx.gen:
            raise RuntimeError("gwb200 UNet1D: another grad-enabled forward of the same (batch, length) ran before this backward; "
                               "the saved activations were overwritten (call backward() first, or use torch.no_grad())")
        flat = torch.zeros(bwd.layout.total, device=d_eps.device, dtype=torch.float32)
        B, _, L = ctx.net.shape
        bwd.backward(ctx.net, d_eps.contiguous().float().reshape(B, L), flat)
        views = bwd.layout.views(flat)
        return (None, None, None, None) + tuple(views[k] for k in ctx.names)


def unet_autograd_forward(model, x: Tensor, t: Tensor, compute_dtype: str) -> Tensor:
    params = [p for _, p in model.named_parameters()]
    return _UNetFunction.apply(model, compute_dtype, x, t, *params)


# ======================================================================================================
# reference helper API (train.py:40-172)
# ======================================================================================================
@torch.no_grad()
def _predict_x0_norm(model, diffusion, x_t: Tensor, cond_stack: Tensor, t: Tensor) -> Tensor:
    """train.py:40-51: one-step x0 estimate with a zero self-conditioning channel."""
    t = t.long()
    net_in = torch.cat([x_t, cond_stack, torch.zeros_like(x_t)], dim=1).contiguous()
    eps_hat = model(net_in, t)
    lib = _cabi.load()
    B, Cx, L = net_in.shape
    ab = diffusion.alpha_bar.to(x_t.device).float().contiguous()
    check(lib.gw_selfcond_x0(ptr(net_in), ptr(eps_hat.contiguous()), ptr(t.contiguous()), ptr(ab), B, Cx, L, _cabi.stream_ptr()),
          "selfcond_x0")
    return net_in[:, Cx - 1:Cx].clone()


def _element_loss(eps_hat: Tensor, eps: Tensor, mask: Tensor, loss_type: str, huber_beta: float) -> Tensor:
    """train.py:53-58 (elementwise; kept as torch ops for reference-style loops -- the fused step uses gw_loss)."""
    if loss_type == "huber":
        el = torch.nn.functional.smooth_l1_loss(eps_hat, eps, reduction="none", beta=huber_beta)
    else:
        el = (eps_hat - eps) ** 2
    return el * mask


@torch.no_grad()
def update_ema(ema_model, model, decay: float) -> None:
    """train.py:73-81."""
    ema_params = dict(ema_model.named_parameters())
    for k, p in model.named_parameters():
        ema_params[k].data.mul_(decay).add_(p.data, alpha=(1.0 - decay))
    for eb, mb in zip(ema_model.buffers(), model.buffers()):
        eb.copy_(mb)


def warmup_cosine_lambda(step: int, warmup_steps: int, total_steps: int, min_lr_scale: float = 0.1) -> float:
    """train.py:85-90."""
    if step < warmup_steps:
        return max(1e-8, float(step + 1) / max(1, warmup_steps))
    progress = (step - warmup_steps) / max(1, (total_steps - warmup_steps))
    progress = min(max(progress, 0.0), 1.0)
    return min_lr_scale + 0.5 * (1 - min_lr_scale) * (1 + math.cos(math.pi * progress))


def make_warmup_cosine_scheduler(optimizer, warmup_steps: int, total_steps: int, min_lr_scale: float = 0.1):
    """train.py:84-91."""
    return torch.optim.lr_scheduler.LambdaLR(
        optimizer, lambda step: warmup_cosine_lambda(step, warmup_steps, total_steps, min_lr_scale))


def _match_batch(a: Tensor, target_bsz: int) -> Tensor:
    """train.py:133-144."""
    if a.shape[0] == target_bsz or a.shape[0] == 0:
        return a
    if target_bsz % a.shape[0] == 0:
        return a.repeat_interleave(target_bsz // a.shape[0], dim=0)
    rep = (target_bsz + a.shape[0] - 1) // a.shape[0]
    return a.repeat_interleave(rep, dim=0)[:target_bsz]


def _sample_timesteps_stratified(bsz: int, t_min: int, t_max: int, device, bins: int = 0) -> Tensor:
    """train.py:147-172; the bucket edges are computed on the host (no per-bucket `.item()` syncs)."""
    b = int(bins) if bins and bins > 0 else int(bsz)
    b = max(1, min(b, bsz))
    edges = torch.linspace(t_min, t_max + 1, b + 1).long().tolist()
    q, r = divmod(bsz, b)
    lo = torch.tensor([edges[i] for i in range(b) for _ in range(q + 1 if i < r else q)], device=device)
    hi = torch.tensor([max(edges[i + 1] - 1, edges[i]) for i in range(b) for _ in range(q + 1 if i < r else q)], device=device)
    t = lo + (torch.rand(lo.numel(), device=device) * (hi - lo + 1).float()).long().clamp_(max=hi - lo)
    return t[torch.randperm(bsz, device=device)].long()


# ======================================================================================================
# fused training step
# ======================================================================================================
class FusedTrainStep:
    """The per-batch body of `train_diffusion` (train.py:320-456) over flat buffers.

    model parameters are re-pointed at views of `self.flat_p` (ParamLayout order), so `model.state_dict()` always shows
    the live weights.  `self.flat_ema` holds the EMA copy (`ema_state_dict()`), `self.flat_m` / `self.flat_v` AdamW state.
    """

    def __init__(self, model, diffusion, B: int, L: int, *, lr: float = 2e-4, weight_decay: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, clip_grad: float = 1.0, ema_decay: Optional[float] = 0.999,
                 loss: str = "huber", huber_beta: float = 0.5, loss_weight_power: float = 0.0, clamp_inputs: float = 10.0,
                 p_uncond: float = 0.2, dropout_y_only: bool = True, t_min: int = 500, warmup_steps: int = 0,
                 total_steps: int = 0, min_lr_scale: float = 0.1, cosine_decay: bool = False, seed: int = 0,
                 compute_dtype: Optional[str] = None, conv_impl: str = "auto", process_group=None, sample0: int = 0,
                 skip_loss_threshold: float = 0.0, clamp_cond_y: bool = False, world: Optional[int] = None,
                 share: Optional["FusedTrainStep"] = None):
        """`skip_loss_threshold` > 0 is `--skip_bad_batches --skip_loss_threshold` (train.py:428-436), decided on the device.
        `clamp_cond_y`: the y channel of the conditioning stack is the clamped one (what train.py:360-369 builds for t_multi > 1).
        `share`: another stepper of the same model (a different batch size / length): parameters, gradients, AdamW moments,
        EMA, the Philox step counter and the applied-step counter are SHARED, so alternating between shapes continues one
        optimisation run (train_diffusion on ragged batches)."""
        import torch.distributed as dist
        self.model, self.diffusion = model, diffusion
        self.B, self.L = B, L
        sp: ModelSpec = model.spec
        self.spec = sp
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("gwb200 FusedTrainStep runs on CUDA (sm_100a) only: no CPU fallback")
        self.device = dev
        self.lib = _cabi.load()
        cd = compute_dtype or model.compute_dtype
        # ---- flat buffers; the module's parameters become views
        shapes = {k: tuple(p.shape) for k, p in model.named_parameters()}
        self.layout = ParamLayout(sp, shapes)
        n = self.layout.total
        if share is not None:
            if share.model is not model:
                raise ValueError("FusedTrainStep(share=...): the steppers must train the same model")
            self.flat_p, self.bucket, self.flat_g = share.flat_p, share.bucket, share.flat_g
            self.flat_m, self.flat_v, self.flat_ema = share.flat_m, share.flat_v, share.flat_ema
            views = self.layout.views(self.flat_p)
        else:
            self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
            views = self.layout.views(self.flat_p)
            with torch.no_grad():
                for k, p in model.named_parameters():
                    views[k].copy_(p.data)
                    p.data = views[k]
            # gradient bucket: n gradients + one slot that carries the batch loss through the all-reduce (see gw_bucket_reset)
            self.bucket = torch.zeros(n + 4, device=dev, dtype=torch.float32)
            self.flat_g = self.bucket[:n]
            self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
            self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
            self.flat_ema = self.flat_p.clone() if ema_decay is not None else None
        self.eng = UNetEngine(views, sp, dtype=cd, conv_impl=conv_impl)
        self.eng.bind_flat(self.flat_p, self.layout)
        self.bwd = BackwardEngine(self.eng, self.layout)
        # ---- hyper-parameters
        self.lr, self.wd, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        self.clip_grad = float(clip_grad)
        self.ema_decay = ema_decay
        self.loss_type = 0 if loss == "huber" else 1
        self.huber_beta, self.lwp = float(huber_beta), float(loss_weight_power)
        self.clamp, self.p_uncond, self.dropout_y_only = float(clamp_inputs), float(p_uncond), bool(dropout_y_only)
        self.t_min, self.T = int(t_min), int(diffusion.T)
        self.warmup_steps, self.total_steps, self.min_lr_scale = int(warmup_steps), int(total_steps), float(min_lr_scale)
        self.use_sched = warmup_steps > 0 or cosine_decay
        self.seed, self.sample0 = int(seed) & (2 ** 64 - 1), int(sample0)
        self.skip_loss_threshold, self.clamp_cond_y = float(skip_loss_threshold), bool(clamp_cond_y)
        self.pg = process_group
        if world is not None:
            self.world = int(world)
        else:
            self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        # ---- step state
        Cx, Cc = sp.in_ch, sp.cond_in_ch
        f32 = dict(device=dev, dtype=torch.float32)
        self.net = torch.zeros(B, Cx, L, **f32)
        self.clean = torch.zeros(B, L, **f32)
        self.cond = torch.zeros(B, max(Cc, 1), L, **f32)
        self.mask = torch.ones(B, L, **f32)
        self.eps_buf = torch.zeros(B, L, **f32)
        self.eps_hat = torch.zeros(B, 1, L, **f32)
        self.d_eps = torch.zeros(B, L, **f32)
        self.t = torch.zeros(B, device=dev, dtype=torch.int64)
        self.drop = torch.zeros(B, **f32)
        self.wt = torch.ones(B, **f32) if self.lwp != 0.0 else None
        self.per_sample = torch.zeros(B, **f32)
        self.loss = torch.zeros(1, **f32)
        self.info = torch.zeros(8, **f32)
        self.hyper = torch.zeros(16, **f32)           # run constants; LR schedule / bias corrections are evaluated on the device
        self._hyper_sent = None
        self.partial = torch.zeros(self.lib.gw_opt_scratch_doubles(), device=dev, dtype=torch.float64)
        if share is not None:
            self.step_ctr, self.opt_state, self._shared = share.step_ctr, share.opt_state, share._shared
        else:
            self.step_ctr = torch.zeros(1, device=dev, dtype=torch.int32)     # Philox draw counter (every attempted step)
            self.opt_state = torch.zeros(4, device=dev, dtype=torch.int32)    # [0] steps applied, [1] batches skipped
            self._shared = {"steps_done": 0}
        ab = diffusion.alpha_bar.to(dev).float().contiguous()
        self.ab, self.sab, self.s1mab = ab, ab.sqrt().contiguous(), (1 - ab).sqrt().contiguous()
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self.overlap_prep = True                     # dgrad weight preparation on a forked stream / graph branch
        import os as _os
        self.capture_allreduce = _os.environ.get("GWB200_GRAPH_ALLREDUCE", "0") == "1"     # world > 1: collective inside the graph
        self._side: Optional[torch.cuda.Stream] = None

    # ------------------------------------------------------------------ host -> device staging
    def prefetch(self, clean_norm: Tensor, cond_stack: Tensor, mask: Optional[Tensor] = None) -> None:
        """Start the H2D copy of the NEXT batch (pinned host tensors) on a side stream into one of two staging sets, so
        it overlaps the step in flight (the reference gets the same overlap from DataLoader pin_memory + non_blocking,
        train.py:184-198, 323-332).  `step(prefetched=True)` consumes it."""
        if not hasattr(self, "_stage"):
            self._stage = [{"clean": torch.empty_like(self.clean), "cond": torch.empty_like(self.cond),
                            "mask": torch.empty_like(self.mask)} for _ in range(2)]
            self._copy_stream = torch.cuda.Stream()
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._free = [torch.cuda.Event(), torch.cuda.Event()]
            for ev in self._free:
                ev.record()
            self._pf, self._cs = 0, 0
        slot = self._pf
        self._pf ^= 1
        B, L = self.B, self.L
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._free[slot])
            sset = self._stage[slot]
            sset["clean"].copy_(clean_norm.reshape(B, L), non_blocking=True)
            if self.spec.cond_in_ch > 0:
                sset["cond"].copy_(cond_stack.reshape(B, self.spec.cond_in_ch, L), non_blocking=True)
            sset["has_mask"] = mask is not None
            if mask is not None:
                sset["mask"].copy_(mask.reshape(B, L), non_blocking=True)
            self._ready[slot].record()

    def _consume_prefetched(self) -> None:
        slot = self._cs
        self._cs ^= 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[slot])
        sset = self._stage[slot]
        self.clean.copy_(sset["clean"])
        if self.spec.cond_in_ch > 0:
            self.cond.copy_(sset["cond"])
        if sset["has_mask"]:
            self.mask.copy_(sset["mask"])
        else:
            self.mask.fill_(1.0)
        self._free[slot].record(cur)

    # ------------------------------------------------------------------ pieces
    def state_dict_ema(self) -> Dict[str, Tensor]:
        return {k: v.clone() for k, v in self.layout.views(self.flat_ema).items()} if self.flat_ema is not None else {}

    def optimizer_state_dict(self) -> Dict:
        """The flat AdamW moments in `torch.optim.AdamW.state_dict()` layout (parameters indexed in `model.parameters()` order),
        which is what the reference stores under 'optimizer_state' (train.py:610)."""
        m, v = self.layout.views(self.flat_m), self.layout.views(self.flat_v)
        names = [k for k, _ in self.model.named_parameters()]
        applied = self.applied_steps()
        state = {i: {"step": torch.tensor(float(applied)), "exp_avg": m[k].clone(), "exp_avg_sq": v[k].clone()}
                 for i, k in enumerate(names)}
        group = {"lr": self.last_lr if applied > 0 else self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": list(range(len(names)))}
        if self.use_sched:
            group["initial_lr"] = self.lr
        return {"state": state, "param_groups": [group]}

    def load_batch(self, clean_norm: Tensor, cond_stack: Tensor, mask: Optional[Tensor] = None) -> None:
        """clean_norm [B,1,L], cond_stack [B,Cc,L] (sigma-normalised, train.py:336-347), mask [B,1,L]; async H2D if pinned."""
        B, L = self.B, self.L
        self.clean.copy_(clean_norm.reshape(B, L), non_blocking=True)
        if self.spec.cond_in_ch > 0:
            self.cond.copy_(cond_stack.reshape(B, self.spec.cond_in_ch, L), non_blocking=True)
        if mask is not None:
            self.mask.copy_(mask.reshape(B, L), non_blocking=True)
        else:
            self.mask.fill_(1.0)                              # no mask = every sample valid (never the previous batch's mask)

    def load_collated(self, clean_raw: Tensor, noisy_raw: Tensor, sigma: Tensor, mask: Optional[Tensor] = None,
                      meta: Optional[Tensor] = None, repeat: int = 1) -> None:
        """A collated loader batch (dataloader.py:248-268) straight into the step's input buffers: sigma-normalisation, the
        [y | metadata] stack and the --t_multi repeat (train.py:336-347, 355-360) in ONE kernel (gw_batch_prepare) instead of
        six eager ops and three staging copies.  clean_raw, noisy_raw, mask [B0, 1, L]; sigma [B0]; meta [B0, Cm, L]; B0 * repeat
        must equal this stepper's batch."""
        dev = self.clean.device
        f = lambda a: a.to(dev, non_blocking=True).float().contiguous()
        clean_raw, noisy_raw, sigma = f(clean_raw), f(noisy_raw), f(sigma).reshape(-1)
        B0, L = int(sigma.numel()), self.L
        Cm = self.spec.cond_in_ch - 1
        if B0 * repeat != self.B or clean_raw.numel() != B0 * L or noisy_raw.numel() != B0 * L:
            raise ValueError(f"load_collated: batch {B0} x repeat {repeat}, length {clean_raw.shape[-1]} vs stepper ({self.B}, {L})")
        if Cm > 0:
            if meta is None or meta.shape[1] != Cm:
                raise ValueError(f"load_collated: the model expects {Cm} metadata channels")
            meta = f(meta)
            if meta.size(-1) != L:                            # train.py:341-343
                meta = torch.nn.functional.interpolate(meta, size=L, mode="linear", align_corners=False).contiguous()
        elif Cm < 0:
            raise ValueError("load_collated needs a conditional model (cond_in_ch >= 1)")
        mask = f(mask) if mask is not None else None
        check(self.lib.gw_batch_prepare(ptr(clean_raw), ptr(noisy_raw), ptr(sigma), ptr(mask), ptr(meta) if Cm > 0 else None, Cm, B0,
                                        L, repeat, ptr(self.clean), ptr(self.cond), ptr(self.mask), _cabi.stream_ptr()),
              "batch_prepare")

    # steps attempted (host count, shared between steppers of one run); the APPLIED count lives on the device
    @property
    def steps_done(self) -> int:
        return self._shared["steps_done"]

    @steps_done.setter
    def steps_done(self, v: int) -> None:
        self._shared["steps_done"] = int(v)

    def applied_steps(self) -> int:
        """Optimisation steps actually applied (skipped batches excluded); one device read."""
        return int(self.opt_state[0])

    def skipped_batches(self) -> int:
        return int(self.opt_state[1])

    @property
    def last_lr(self) -> float:
        """Learning rate of the last step (device read)."""
        return float(self.info[3])

    def _set_hyper(self) -> None:
        """Upload the run constants when they changed (normally once): nothing here depends on the step number."""
        h = (self.lr, 0.0, 0.0, self.ema_decay if self.ema_decay is not None else -1.0, self.wd, self.clip_grad,
             1.0 / self.world, self.skip_loss_threshold, float(self.warmup_steps), float(self.total_steps), self.min_lr_scale,
             1.0 if self.use_sched else 0.0, 0.0, 0.0, 0.0, 0.0)
        if h != self._hyper_sent:
            self.hyper.copy_(torch.tensor(h, dtype=torch.float32))
            self._hyper_sent = h

    def _enqueue(self, selfcond: bool, draws: bool, philox: bool) -> None:
        """All kernels of one step on the current stream (capturable: no host reads, no allocation)."""
        lib, eng, sp = self.lib, self.eng, self.spec
        B, L, Cx, Cc = self.B, self.L, sp.in_ch, sp.cond_in_ch
        st = _cabi.stream_ptr()
        if draws:
            check(lib.gw_train_draws(self.seed, ptr(self.step_ctr), self.sample0, B, self.t_min, self.T, self.p_uncond,
                                     ptr(self.t), ptr(self.drop), st), "train_draws")
            eng.launches += 1
        # fork: the dgrad weights depend on the parameters only -> a second stream (a parallel branch of the captured graph)
        # prepares them while the forward pass runs; joined before the backward pass
        cur = torch.cuda.current_stream()
        if self.overlap_prep:
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.bwd.prepare_dgrad(eng.workspace(B, L, True))
        use_drop = self.p_uncond > 0.0
        y_only = self.dropout_y_only and Cc > 1
        check(lib.gw_train_pack(ptr(self.clean), ptr(self.cond) if Cc > 0 else None, Cc, ptr(self.t),
                                ptr(self.drop) if use_drop else None, ptr(self.sab), ptr(self.s1mab), ptr(self.eps_buf),
                                1 if philox else 0, self.seed, self.sample0, ptr(self.step_ctr), self.clamp,
                                1 if ((use_drop and y_only) or self.clamp_cond_y) else 0, 0 if y_only else 1, ptr(self.net), B,
                                Cx, L, st), "train_pack")
        eng.launches += 1
        if selfcond and sp.use_selfcond:                      # train.py:401-403: extra no-grad forward, zero self-cond
            self.bwd.forward(self.net, self.t, self.eps_hat)
            check(lib.gw_selfcond_x0(ptr(self.net), ptr(self.eps_hat), ptr(self.t), ptr(self.ab), B, Cx, L, st), "selfcond_x0")
            eng.launches += 1
        self.bwd.forward(self.net, self.t, self.eps_hat)
        if self.wt is not None:
            check(lib.gw_loss_weight(ptr(self.t), ptr(self.ab), self.lwp, ptr(self.wt), B, st), "loss_weight")
            eng.launches += 1
        check(lib.gw_loss(ptr(self.eps_hat), ptr(self.eps_buf), ptr(self.mask), ptr(self.wt), B, L, self.loss_type,
                          self.huber_beta, 1.0, ptr(self.per_sample), ptr(self.loss), ptr(self.d_eps), st), "loss")
        eng.launches += 2
        check(lib.gw_bucket_reset(ptr(self.bucket), self.layout.total, ptr(self.loss), st), "bucket_reset")
        eng.launches += 1
        if self.overlap_prep:
            cur.wait_stream(self._side)
        self.bwd._prepped = self.overlap_prep
        try:
            self.bwd.backward(self.net, self.d_eps, self.flat_g)
        finally:
            self.bwd._prepped = False

    def _enqueue_update(self) -> None:
        lib, st = self.lib, _cabi.stream_ptr()
        n = self.layout.total
        check(lib.gw_grad_sumsq(ptr(self.flat_g), n, ptr(self.partial), st), "grad_sumsq")
        check(lib.gw_adamw_ema(ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_m), ptr(self.flat_v), ptr(self.flat_ema), n,
                               ptr(self.partial), ptr(self.hyper), None, 1, ptr(self.opt_state), float(self.betas[0]),
                               float(self.betas[1]), self.eps, ptr(self.info), st), "adamw_ema")
        self.eng.refresh()                                    # re-pack the bf16 conv weights from the updated fp32 master
        check(lib.gw_train_advance(ptr(self.step_ctr), ptr(self.opt_state), ptr(self.info), st), "train_advance")
        self.eng.launches += 3

    def _allreduce(self) -> None:
        """Sum of the gradient bucket (+ the loss slot) over the ranks; the 1/world scale is applied in gw_adamw_ema.  NCCL
        collectives are capturable, so inside `_capture` this becomes a node of the step's CUDA graph."""
        if self.world > 1:
            from .parallel import allreduce_flat_
            allreduce_flat_(self.bucket[: self.layout.total + 1], self.pg)

    # ------------------------------------------------------------------ public
    def step(self, *, selfcond: bool = False, t: Optional[Tensor] = None, eps: Optional[Tensor] = None,
             drop: Optional[Tensor] = None, use_graph: bool = True, prefetched: bool = False) -> None:
        """One optimisation step on the batch last given to `load_batch`.

        `t` / `eps` / `drop` inject the draws the reference takes from the global RNG (parity tests); left None they come
        from the on-device Philox streams keyed on (seed, global sample index, step).  `selfcond` is the per-batch coin of
        train.py:401 (drawn by the caller from a host RNG so that no device->host sync is needed)."""
        draws = t is None
        philox = eps is None
        if prefetched:
            self._consume_prefetched()
        if t is not None:
            self.t.copy_(t.reshape(-1).long(), non_blocking=True)
        if eps is not None:
            self.eps_buf.copy_(eps.reshape(self.B, self.L), non_blocking=True)
        if drop is not None:
            self.drop.copy_(drop.reshape(-1).float(), non_blocking=True)
        elif not draws:
            self.drop.zero_()
        self._set_hyper()
        if self._shared.get("owner") is not self:             # another stepper of this run updated the shared parameters:
            if self._shared.get("owner") is not None:         # my packed bf16 conv weights are stale
                self.eng.refresh()
            self._shared["owner"] = self
        if not use_graph:
            self._enqueue(selfcond, draws, philox)
            self._allreduce()
            self._enqueue_update()
        else:
            key = (bool(selfcond), draws, philox, self.p_uncond, self.t_min)
            gs = self._graphs.get(key)
            if gs is None:
                gs = self._capture(key)
            if len(gs) == 1:
                gs[0].replay()
            else:                                             # world > 1 (default): the collective runs between two graphs
                gs[0].replay()
                self._allreduce()
                gs[1].replay()
        self.steps_done += 1
        # the flat buffer changed under the module's parameter views (no torch version bump): engines that UNet1D.engine()
        # hands out (model(x, t), ddim_sample(model, ...)) must re-pack their bf16 weights / FiLM tables on next use
        self.model._versions = None

    def _capture(self, key):
        """world == 1: ONE graph per step flavour: [draws, pack, (self-cond fwd), fwd, loss, backward, norm, clip+AdamW+EMA,
        re-pack, advance] -- a step is a single graph launch.  world > 1: two graphs with the NCCL all-reduce of the gradient
        bucket between them on the same stream (the configuration measured at 2 / 8 GPUs); `capture_allreduce = True` captures
        the collective into the single graph as well (NCCL is capturable; opt-in)."""
        selfcond, draws, philox = key[:3]
        one = self.world == 1 or self.capture_allreduce
        saved = [b.clone() for b in (self.flat_p, self.flat_m, self.flat_v, self.step_ctr, self.opt_state)]
        saved_ema = self.flat_ema.clone() if self.flat_ema is not None else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                            # warm-up: lazy packing, cudaFuncSetAttribute, NCCL channels
            self._enqueue(selfcond, draws, philox)
            self._allreduce()
            self._enqueue_update()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for b, sv in zip((self.flat_p, self.flat_m, self.flat_v, self.step_ctr, self.opt_state), saved):
            b.copy_(sv)
        if saved_ema is not None:
            self.flat_ema.copy_(saved_ema)
        self.eng.refresh()
        if one:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(selfcond, draws, philox)
                self._allreduce()
                self._enqueue_update()
            self._graphs[key] = (g,)
        else:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._enqueue(selfcond, draws, philox)
            with torch.cuda.graph(g2):
                self._enqueue_update()
            self._graphs[key] = (g1, g2)
        return self._graphs[key]


# ======================================================================================================
# train_diffusion (train.py:174-630): host loop around FusedTrainStep
# ======================================================================================================
def _compute_meta_scale(h5_file: str) -> dict:
    """train.py:105-130: dataset-adaptive label scales (95th percentile of the masses / mass ratio)."""
    import numpy as np
    from .dataloader import open_h5
    scale = {"M": 80.0, "q": 10.0}
    try:
        f = open_h5(h5_file)
        try:
            def p95(name):
                if name in f:
                    arr = np.array(f[name][...], dtype=np.float64)
                    if arr.size:
                        return float(np.nanpercentile(arr, 95))
                return None
            Ms = [x for x in (p95("mass1"), p95("mass2"), p95("chirp_mass")) if (x is not None and np.isfinite(x) and x > 0)]
            if Ms:
                scale["M"] = float(max(Ms))
            q_p = p95("q")
            if (q_p is not None) and np.isfinite(q_p) and q_p > 0:
                scale["q"] = float(q_p)
        finally:
            f.close()
    except Exception as e:
        print(f"[train] meta_scale computation failed; using defaults {scale} ({e})")
    return scale


def train_diffusion(args, loader=None):
    """Reference entry point `train_diffusion(args)` (train.py:174-630).  With `loader=None` the data loader is built from
    `args.data` exactly as the reference does (train.py:178-198: meta scale, then `make_dataloader`), through
    `dataloader.BatchLoader` (pinned double-buffered staging, whitening / sigma on the GPU).  `loader` may instead be any
    iterable of (clean, noisy, sigma, mask[, meta]) batches shaped like `dataloader.pad_collate` output."""
    import random
    from .models import CustomDiffusion, UNet1D
    meta_scale = None
    if loader is None:
        from .dataloader import make_dataloader, resolve_h5_path
        h5_path = resolve_h5_path(args.data)
        meta_scale = _compute_meta_scale(h5_path)
        loader = make_dataloader(h5_path=args.data, batch_size=args.batch_size, shuffle=True,
                                 num_workers=getattr(args, "num_workers", 0), pin_memory=True,
                                 whiten=getattr(args, "whiten", False), whiten_mode=getattr(args, "whiten_mode", "auto"),
                                 sigma_mode=getattr(args, "sigma_mode", "std"), sigma_fixed=getattr(args, "sigma_fixed", 1.0),
                                 include_metadata=True, mass_scale=float(meta_scale.get("M", 80.0)), device=args.device)
        if len(loader.dataset) == 0:
            raise RuntimeError("Empty dataset")
    if getattr(args, "seed", None) is not None:
        random.seed(args.seed)
        torch.manual_seed(args.seed)
    device = torch.device(args.device)
    peek = next(iter(loader))
    C_meta = int(peek[4].shape[1]) if len(peek) == 5 else 0
    cond_in_ch = 1 + C_meta
    in_ch = 1 + cond_in_ch + 1
    L = int(peek[0].shape[-1])
    K = max(1, int(getattr(args, "t_multi", 1)))
    B = int(peek[0].shape[0]) * K
    model = UNet1D(in_ch=in_ch, base_ch=args.base_ch, time_dim=args.time_dim, depth=args.depth,
                   t_embed_max_time=max(0, args.T - 1), cond_in_ch=cond_in_ch, use_selfcond=True,
                   compute_dtype="bf16" if getattr(args, "amp", False) else "fp32").to(device)
    diffusion = CustomDiffusion(T=args.T, device=device)
    if getattr(args, "init_from", None):
        ckpt = torch.load(args.init_from, map_location=device)
        model.load_state_dict(ckpt.get("model_ema_state", ckpt.get("model_state")), strict=True)
    total_steps = len(loader) * args.epochs
    stepper, first = None, None
    steppers: Dict[tuple, FusedTrainStep] = {}
    max_shapes = int(getattr(args, "max_cached_shapes", 4))
    rng = random.Random(getattr(args, "seed", 0) or 0)
    history = []
    for epoch in range(1, args.epochs + 1):
        forced = epoch <= getattr(args, "force_cond_epochs", 0)
        p_uncond = 0.0 if forced else args.p_uncond
        p_selfcond = 0.0 if forced else args.p_selfcond
        t_min = int(max(0, min(args.T - 1, int(args.t_min_frac * args.T))))
        for batch in loader:
            clean_raw, noisy_raw, sigma, mask = batch[:4]
            meta = batch[4] if (len(batch) == 5 and C_meta > 0) else None
            key = (int(clean_raw.shape[0]) * K, int(clean_raw.shape[-1]))
            stepper = steppers.get(key)
            if stepper is None:
                # pad_collate pads to the per-batch maximum (dataloader.py:248-268), so (B, L) may change from batch to batch:
                # one stepper (workspace + captured graphs) per shape, all SHARING parameters, gradient bucket, AdamW moments,
                # EMA, the Philox draw counter and the applied-step counter -- the run continues, it does not restart
                if len(steppers) >= max_shapes:
                    steppers.pop(next(iter(steppers)))          # oldest shape: its workspace / graphs are released
                thr = float(getattr(args, "skip_loss_threshold", 0.0)) if getattr(args, "skip_bad_batches", False) else 0.0
                stepper = FusedTrainStep(model, diffusion, key[0], key[1], lr=args.lr,
                                         weight_decay=args.weight_decay, clip_grad=args.clip_grad,
                                         ema_decay=args.ema_decay if args.ema else None, loss=args.loss,
                                         huber_beta=args.huber_beta, loss_weight_power=args.loss_weight_power,
                                         clamp_inputs=args.clamp_inputs, p_uncond=p_uncond,
                                         dropout_y_only=args.dropout_y_only, t_min=t_min, warmup_steps=args.warmup_steps,
                                         total_steps=total_steps, min_lr_scale=args.min_lr_scale,
                                         cosine_decay=args.cosine_decay, seed=getattr(args, "seed", 0) or 0,
                                         skip_loss_threshold=thr, clamp_cond_y=(K > 1 and args.clamp_inputs > 0), share=first)
                steppers[key] = stepper
                if first is None:
                    first = stepper
            stepper.p_uncond, stepper.t_min = p_uncond, t_min
            stepper.load_collated(clean_raw, noisy_raw, sigma, mask, meta, repeat=K)   # sigma-normalise + stack + repeat: one kernel
            t_inj = None
            if getattr(args, "t_cover", "rand") == "strat":
                t_inj = _sample_timesteps_stratified(stepper.B, t_min, args.T - 1, device, bins=getattr(args, "t_bins", 0))
            stepper.step(selfcond=(p_selfcond > 0.0 and rng.random() < p_selfcond), t=t_inj,
                         drop=(torch.rand(stepper.B, device=device) < p_uncond).float() if t_inj is not None else None)
            history.append(stepper.loss.clone())              # device scalar; read back once per epoch below
        history = [float(h) if isinstance(h, Tensor) else h for h in history]
    # checkpoint payload of train.py:606-630 (same keys; 'model_ema_state' only with --ema)
    payload = {"model_state": {k: v.detach().clone() for k, v in model.state_dict().items()},
               "optimizer_state": stepper.optimizer_state_dict() if stepper is not None else {},
               "args": {**vars(args), "conditional": True, "in_ch": in_ch, "cond_in_ch": cond_in_ch, "meta_enabled": C_meta > 0,
                        "meta_channels": C_meta,
                        "conditioning": "concat[y + meta]+selfcond" if C_meta > 0 else "concat[y]+selfcond",
                        "whiten": getattr(args, "whiten", False), "whiten_mode": getattr(args, "whiten_mode", "auto"),
                        "sigma_mode": getattr(args, "sigma_mode", "std"), "dropout_y_only": bool(args.dropout_y_only),
                        "meta_scale": meta_scale if meta_scale is not None else getattr(args, "meta_scale", {"M": 80.0, "q": 10.0})},
               "epoch": args.epochs}
    if getattr(args, "ema", False) and stepper is not None:
        payload["model_ema_state"] = stepper.state_dict_ema()
    if getattr(args, "model_dir", None):                       # train.py:17-27, 607: <model_dir>/latest_model/model_diffusion.pth
        import os
        out_dir = os.path.join(args.model_dir, "latest_model")
        os.makedirs(out_dir, exist_ok=True)
        torch.save(payload, os.path.join(out_dir, "model_diffusion.pth"))
    return {"model": model, "stepper": stepper, "losses": history, "checkpoint": payload}
