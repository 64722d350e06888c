"""Tensor-level wrappers over the C ABI: torch is used for device memory and the current stream only.

Every function launches hand-written sm_100a kernels from libviditq_b200.so; none has a PyTorch/CPU fallback.
"""
import functools
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import VQ_EPI_BIAS, VQ_EPI_GELU_TANH, VQ_EPI_GATE_RESIDUAL  # noqa: F401

_launches = 0          # kernels launched through this module (bench.py reports it as gpu_launches)
_status = {}           # device index -> uint32 status word


def launch_count():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _nvtx(fn):
    """NVTX range around every kernel wrapper when VQ_NVTX=1 (nsys / ncu --nvtx timelines name the fused op, not only the
    kernel); a no-op otherwise."""
    if os.environ.get("VQ_NVTX", "0") != "1":
        return fn

    @functools.wraps(fn)
    def inner(*a, **k):
        torch.cuda.nvtx.range_push("viditq_b200." + fn.__name__)
        try:
            return fn(*a, **k)
        finally:
            torch.cuda.nvtx.range_pop()
    return inner


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda_f16(t, name):
    if not (t.is_cuda and t.dtype == torch.float16 and t.is_contiguous()):
        raise _lib.VqError(f"{name}: expected a contiguous CUDA fp16 tensor, got {t.dtype} {t.device} "
                           f"contiguous={t.is_contiguous()} (viditq_b200 has no CPU path)")


def status_word(device=None):
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev not in _status:
        _status[dev] = torch.zeros(1, dtype=torch.int32, device=f"cuda:{dev}")
    return _status[dev]


@_nvtx
def cfg_ddim_step(out_cond, out_uncond, x, coef, cfg_scale, ptqd_k=0.0, out=None):
    """Fused CFG combine + DDIM (eta = 0) update. out_cond / out_uncond: fp32 CUDA [n, 2c, ...]; x: fp32 [n, c, ...];
    coef: fp32 CUDA [4] (SpacedDDIM.coefficients). Returns the next latent (fp32, x's shape)."""
    for t, name in ((out_cond, "out_cond"), (out_uncond, "out_uncond"), (x, "x"), (coef, "coef")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise _lib.VqError(f"cfg_ddim_step: {name} must be a contiguous CUDA fp32 tensor (no CPU fallback)")
    n, c = x.shape[0], x.shape[1]
    inner = x.numel() // (n * c)
    if out_cond.shape != out_uncond.shape or out_cond.shape[0] != n or out_cond.numel() != n * out_cond.shape[1] * inner:
        raise _lib.VqError(f"cfg_ddim_step: shapes {tuple(out_cond.shape)} / {tuple(x.shape)} do not match")
    if out is None:
        out = torch.empty_like(x)
    rc = _lib.lib().vq_cfg_ddim_step(_ptr(out_cond), _ptr(out_uncond), _ptr(x), _ptr(coef), float(cfg_scale),
                                     float(ptqd_k), n, out_cond.shape[1], c, inner, _ptr(out), _stream())
    _lib.check(rc, "vq_cfg_ddim_step")
    _count()
    return out


@_nvtx
def patch_embed(latent, weight, bias, pos, patch_hw, T=None):
    """Fused patchify + position embedding. latent: fp32 CUDA [B, Cin, T, H, W] (or [B, Cin, H, W]); weight: fp16 conv
    weight [C, Cin, (1,) ph, pw]; bias fp16 [C] or None; pos fp16 [S, C] or None -> fp16 [B, T*S, C]."""
    if not (latent.is_cuda and latent.dtype == torch.float32 and latent.is_contiguous()):
        raise _lib.VqError("patch_embed: latent must be a contiguous CUDA fp32 tensor (no CPU fallback)")
    if latent.dim() == 4:
        latent = latent.unsqueeze(2)
    B, Cin, T_, Hh, Ww = latent.shape
    ph, pw = patch_hw
    C = weight.shape[0]
    w2 = weight.reshape(C, -1)
    _need_cuda_f16(w2, "weight")
    if w2.shape[1] != Cin * ph * pw:
        raise _lib.VqError(f"patch_embed: weight {tuple(weight.shape)} is not a depth-1 ({ph}, {pw}) patch kernel")
    if bias is not None:
        _need_cuda_f16(bias, "bias")
    S = (Hh // ph) * (Ww // pw)
    if pos is not None:
        _need_cuda_f16(pos, "pos")
        if pos.numel() != S * C:
            raise _lib.VqError(f"patch_embed: pos has {pos.numel()} elements, expected {S * C}")
    out = torch.empty((B, T_ * S, C), dtype=torch.float16, device=latent.device)
    rc = _lib.lib().vq_patch_embed(_ptr(latent), _ptr(w2), _ptr(bias), _ptr(pos), B, Cin, T_, Hh, Ww, ph, pw, C, _ptr(out),
                                   _stream())
    _lib.check(rc, "vq_patch_embed")
    _count()
    return out


def check_status(device=None):
    """Poll the sticky device status word (synchronises; call outside the hot loop). Raises on the reference's
    degenerate-eps quirk (base_quantizer.py:220-223), whose fp16 result is non-finite garbage in the reference."""
    word = int(status_word(device).item())
    if word & _lib.VQ_STATUS_EPS_DEGENERATE:
        status_word(device).zero_()
        raise _lib.VqError("dynamic act quant: a token row had delta < 1e-6 (reference quirk Q4: delta.fill_(1e-6) "
                           "for ALL rows; in fp16 that yields inf/NaN) — refusing to continue")
    return word


@dataclass
class PreparedWeight:
    """u8 codes [N,K] + per-channel {c1, zw, dw, bias} records (16 B each) of one (layer, n_bits, timerange)."""
    codes: torch.Tensor
    col: torch.Tensor
    N: int
    K: int
    n_bits: int
    smooth: Optional[torch.Tensor] = None   # fp16 [K] smooth-quant channel scale folded into the codes (or None)
    packed: Optional[torch.Tensor] = None   # u8 [N, K/2]: the same codes two per byte (n_bits <= 4; vq_linear_w4a8)


@dataclass
class ActCodes:
    """u8 codes [G*rows,K], fp16 delta/zp [rows] (token statistics pooled over G), i32 rowsum [G*rows]."""
    codes: torch.Tensor
    delta: torch.Tensor
    zp: torch.Tensor
    rowsum: torch.Tensor
    G: int
    rows: int
    K: int
    pw: Optional[PreparedWeight] = None   # set by QuantLayer.quantize_input: the weight these codes were scaled for


@_nvtx
def prep_weight(w, delta, zp, n_bits=8, smooth=None, bias=None) -> PreparedWeight:
    _need_cuda_f16(w, "weight")
    N, K = w.shape
    delta = delta.reshape(-1).to(torch.float16).contiguous()
    zp = zp.reshape(-1).to(torch.float16).contiguous()
    if delta.numel() != N or zp.numel() != N:
        raise _lib.VqError("prep_weight: per-output-channel delta/zero_point expected")
    if smooth is not None:
        smooth = smooth.reshape(-1).to(torch.float16).contiguous()
    if bias is not None:
        bias = bias.reshape(-1).to(torch.float16).contiguous()
    codes = torch.empty((N, K), dtype=torch.uint8, device=w.device)
    col = torch.empty((N, 4), dtype=torch.int32, device=w.device)
    rc = _lib.lib().vq_prep_weight(_ptr(w), _ptr(delta), _ptr(zp), _ptr(smooth), _ptr(bias), N, K, n_bits,
                                   _ptr(codes), _ptr(col), _stream())
    _lib.check(rc, "vq_prep_weight")
    _count()
    return PreparedWeight(codes, col, N, K, n_bits)


@_nvtx
def pack_u4(w: PreparedWeight) -> PreparedWeight:
    """Attach the packed-INT4 form of a prepared weight whose codes are < 16 (n_bits <= 4): two codes per byte."""
    if w.n_bits > 4:
        raise _lib.VqError(f"pack_u4: {w.n_bits}-bit codes do not fit a nibble")
    if w.packed is None:
        packed = torch.empty((w.N, w.K // 2), dtype=torch.uint8, device=w.codes.device)
        rc = _lib.lib().vq_pack_u4(_ptr(w.codes), w.N, w.K, _ptr(packed), _stream())
        _lib.check(rc, "vq_pack_u4")
        _count()
        w.packed = packed
    return w


def _alloc_act(G, rows, K, device):
    return ActCodes(torch.empty((G * rows, K), dtype=torch.uint8, device=device),
                    torch.empty(rows, dtype=torch.float16, device=device),
                    torch.empty(rows, dtype=torch.float16, device=device),
                    torch.empty(G * rows, dtype=torch.int32, device=device), G, rows, K)


@_nvtx
def act_quant(x, n_bits=8, smooth=None, out: Optional[ActCodes] = None, gelu=False) -> ActCodes:
    """x: fp16 [G, rows, K] (reference layout [BS, n_token, C]); statistics per token pooled over G.
    gelu=True quantises gelu_tanh(x) instead (the Mlp activation fused in front of fc2's quantiser)."""
    _need_cuda_f16(x, "x")
    G, rows, K = x.shape
    a = out if out is not None else _alloc_act(G, rows, K, x.device)
    if smooth is not None:
        _need_cuda_f16(smooth, "smooth")
    fn = _lib.lib().vq_gelu_act_quant if gelu else _lib.lib().vq_act_quant
    rc = fn(_ptr(x), G, rows, K, rows * K, K, _ptr(smooth), n_bits, _ptr(a.codes), _ptr(a.delta),
            _ptr(a.zp), _ptr(a.rowsum), _ptr(status_word(x.device)), _stream())
    _lib.check(rc, "vq_gelu_act_quant" if gelu else "vq_act_quant")
    _count()
    return a


@_nvtx
def add_act_quant(x, addv, rows_per_add, n_bits=8, smooth=None) -> ActCodes:
    """Quantise h(x + addv[(r // rows_per_add) % len(addv)]) per token: x fp16 [G, rows, K], addv fp16 [period, K]."""
    _need_cuda_f16(x, "x")
    _need_cuda_f16(addv, "addv")
    G, rows, K = x.shape
    if addv.dim() != 2 or addv.shape[1] != K:
        raise _lib.VqError(f"add_act_quant: addv {tuple(addv.shape)} does not match K={K}")
    a = _alloc_act(G, rows, K, x.device)
    if smooth is not None:
        _need_cuda_f16(smooth, "smooth")
    rc = _lib.lib().vq_add_act_quant(_ptr(x), _ptr(addv), int(rows_per_add), addv.shape[0], G, rows, K, _ptr(smooth), n_bits,
                                     _ptr(a.codes), _ptr(a.delta), _ptr(a.zp), _ptr(a.rowsum), _ptr(status_word(x.device)),
                                     _stream())
    _lib.check(rc, "vq_add_act_quant")
    _count()
    return a


@_nvtx
def col_absmax(x, gelu=False):
    """input.abs().max(dim=-2)[0] of the smooth-quant statistics (quant_layer.py:116,119): x fp16 [G, n, K] -> fp16 [G, K]
    (exact).  gelu=True: maxima of h(gelu_tanh(x))."""
    _need_cuda_f16(x, "x")
    G, n, K = x.shape
    bits = torch.zeros((G, K), dtype=torch.int32, device=x.device)
    rc = _lib.lib().vq_col_absmax(_ptr(x), G, n, K, 1 if gelu else 0, _ptr(bits), _stream())
    _lib.check(rc, "vq_col_absmax")
    _count()
    return bits.to(torch.int16).view(torch.float16)


@_nvtx
def pack_rows(a: ActCodes, dims, strides):
    """ActCodes (G == 1) -> u8 [rows, K + 16] exchange rows (codes + {delta, zp, rowsum} tail), row (i0, i1, i2, i3) of the
    4-D source order `dims` written at position sum(i * stride)."""
    rows, K = a.codes.shape[0], a.K
    out = torch.empty((rows, K + 16), dtype=torch.uint8, device=a.codes.device)
    rc = _lib.lib().vq_row_pack(_ptr(a.codes), _ptr(out), _ptr(a.delta), _ptr(a.zp), _ptr(a.rowsum), rows, K, dims[1], dims[2],
                                dims[3], strides[0], strides[1], strides[2], strides[3], 0, _stream())
    _lib.check(rc, "vq_row_pack")
    _count()
    return out


@_nvtx
def unpack_rows(buf, K, dims, strides) -> ActCodes:
    """Inverse of pack_rows on a received buffer: u8 [rows, K + 16] in source order `dims` -> ActCodes in the permuted order."""
    rows = buf.shape[0]
    a = _alloc_act(1, rows, K, buf.device)
    rc = _lib.lib().vq_row_pack(_ptr(buf), _ptr(a.codes), _ptr(a.delta), _ptr(a.zp), _ptr(a.rowsum), rows, K, dims[1], dims[2],
                                dims[3], strides[0], strides[1], strides[2], strides[3], 1, _stream())
    _lib.check(rc, "vq_row_pack")
    _count()
    return a


@_nvtx
def act_quant_static(x, delta, zp, n_bits=8, smooth=None, gelu=False, ln=None, rows_per_mod=None) -> ActCodes:
    """Static (calibrated) activation scales: x fp16 [..., K]; delta / zp: fp16 CUDA [period] (1 = per-tensor, the
    w8a8_naive.yaml case; rows = static per-token). Row m uses index m % period.
    gelu=True: the codes of h(gelu_tanh(x)) (vq_gelu_act_quant_static); ln=(shift, scale) fp16 [M / rows_per_mod, K]: the
    codes of LayerNorm + t2i_modulate of x (vq_ln_modulate_act_quant_static; rows_per_mod defaults to x.shape[-2]) — the
    fused schedule's one-pass forms of the static quantiser."""
    _need_cuda_f16(x, "x")
    _need_cuda_f16(delta, "delta")
    _need_cuda_f16(zp, "zp")
    K = x.shape[-1]
    M = x.numel() // K
    period = delta.numel()
    if zp.numel() != period or M % period != 0:
        raise _lib.VqError(f"act_quant_static: {M} rows are not a multiple of the {period} (delta, zp) pairs")
    if gelu and ln is not None:
        raise _lib.VqError("act_quant_static: gelu and ln are exclusive")
    codes = torch.empty((M, K), dtype=torch.uint8, device=x.device)
    rowsum = torch.empty(M, dtype=torch.int32, device=x.device)
    if smooth is not None:
        _need_cuda_f16(smooth, "smooth")
    L = _lib.lib()
    if ln is not None:
        shift, scale = ln
        _need_cuda_f16(shift, "shift")
        _need_cuda_f16(scale, "scale")
        rpm = x.shape[-2] if rows_per_mod is None else int(rows_per_mod)
        if shift.numel() != (M // rpm) * K or scale.numel() != shift.numel() or not x.is_contiguous():
            raise _lib.VqError(f"act_quant_static: shift/scale of {shift.numel()} elements do not match {M} rows / {rpm}")
        rc = L.vq_ln_modulate_act_quant_static(_ptr(x), _ptr(shift), _ptr(scale), M, K, rpm, _ptr(delta), _ptr(zp), period,
                                               _ptr(smooth), n_bits, _ptr(codes), _ptr(rowsum), _stream())
        _lib.check(rc, "vq_ln_modulate_act_quant_static")
    else:
        fn = L.vq_gelu_act_quant_static if gelu else L.vq_act_quant_static
        rc = fn(_ptr(x), M, K, K, _ptr(delta), _ptr(zp), period, _ptr(smooth), n_bits, _ptr(codes), _ptr(rowsum), _stream())
        _lib.check(rc, "vq_gelu_act_quant_static" if gelu else "vq_act_quant_static")
    _count()
    return ActCodes(codes, delta.reshape(-1), zp.reshape(-1), rowsum, M // period, period, K)


@_nvtx
def act_quant_heads(x, G, rows, S, n_bits=8, out: Optional[ActCodes] = None) -> ActCodes:
    """x: fp16 head-major attention output [G * rows / S, H, S, 72] (contiguous). Quantises the token-major view
    [G, rows, H*72] without materialising it."""
    _need_cuda_f16(x, "x")
    n, H, S_, D = x.shape
    if S_ != S or n * S != G * rows:
        raise _lib.VqError(f"act_quant_heads: shape {tuple(x.shape)} inconsistent with G={G} rows={rows} S={S}")
    a = out if out is not None else _alloc_act(G, rows, H * D, x.device)
    rc = _lib.lib().vq_act_quant_heads(_ptr(x), G, rows, H, S, D, n_bits, _ptr(a.codes), _ptr(a.delta), _ptr(a.zp),
                                       _ptr(a.rowsum), _ptr(status_word(x.device)), _stream())
    _lib.check(rc, "vq_act_quant_heads")
    _count()
    return a


@_nvtx
def ln_modulate_act_quant(x, shift, scale, n_bits=8, want_y=False, out: Optional[ActCodes] = None, smooth=None,
                          rows_per_mod=None):
    """x: fp16 [G, rows, K]; shift/scale: fp16 [G * rows / rows_per_mod, K] (default one per batch entry); smooth: fp16
    [K] or None. rows_per_mod < rows (G must be 1): samples stacked along the rows, un-pooled statistics (cfg_split).
    Returns (ActCodes, y or None)."""
    _need_cuda_f16(x, "x")
    _need_cuda_f16(shift, "shift")
    _need_cuda_f16(scale, "scale")
    G, rows, K = x.shape
    rpm = rows if rows_per_mod is None else int(rows_per_mod)
    if shift.numel() != (G * rows // rpm) * K or scale.numel() != shift.numel() or (rpm != rows and G != 1):
        raise _lib.VqError(f"ln_modulate_act_quant: shift/scale of {shift.numel()} elements do not match G={G} "
                           f"rows={rows} rows_per_mod={rpm} K={K}")
    a = out if out is not None else _alloc_act(G, rows, K, x.device)
    y = torch.empty_like(x) if want_y else None
    if smooth is not None:
        _need_cuda_f16(smooth, "smooth")
    rc = _lib.lib().vq_ln_modulate_act_quant(_ptr(x), _ptr(shift), _ptr(scale), _ptr(smooth), G, rows, K, rpm, n_bits, _ptr(y),
                                             _ptr(a.codes), _ptr(a.delta), _ptr(a.zp), _ptr(a.rowsum),
                                             _ptr(status_word(x.device)), _stream())
    _lib.check(rc, "vq_ln_modulate_act_quant")
    _count()
    return a, y


@_nvtx
def gemm_w8a8(a: ActCodes, w: PreparedWeight, epi=VQ_EPI_BIAS, res=None, gate=None, rows_per_gate=0, out=None, ldo=None):
    """out[M,N] fp16 = epilogue(dequant(a.codes @ w.codes^T)); M = G*rows. `out` may be a column slice of a wider
    row-major tensor (pass its row pitch as ldo)."""
    M = a.G * a.rows
    if a.K != w.K:
        raise _lib.VqError(f"gemm_w8a8: K mismatch {a.K} vs {w.K}")
    if out is None:
        out = torch.empty((M, w.N), dtype=torch.float16, device=a.codes.device)
    if epi == VQ_EPI_GATE_RESIDUAL:
        _need_cuda_f16(res, "res")
        _need_cuda_f16(gate, "gate")
    rc = _lib.lib().vq_gemm_w8a8(_ptr(a.codes), _ptr(a.delta), _ptr(a.zp), _ptr(a.rowsum), a.rows, _ptr(w.codes),
                                 _ptr(w.col), M, w.N, w.K, epi, _ptr(res), w.N, _ptr(gate), rows_per_gate, _ptr(out),
                                 w.N if ldo is None else ldo, _stream())
    _lib.check(rc, "vq_gemm_w8a8")
    _count()
    return out


_workspaces = {}       # (device index, stream handle) -> uint8 scratch of vq_linear_w8a8's two-launch path


def _workspace(nbytes, device):
    key = (device.index, _stream())
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _workspaces[key] = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
    return buf


def set_linear_fused_policy(mode=0, max_m=0):
    """Schedule behind linear_w8a8: mode 0 quantise pass + GEMM (default), 1 the panel-resident fused kernel on every
    supported shape, -1 the same up to max_m rows, 2 the OVERLAPPED schedule (persistent GEMM with quantiser warpgroups
    running ahead of its MMAs; K = 1152, un-pooled statistics, >= 1024 rows).  Callers restore explicitly."""
    _lib.check(_lib.lib().vq_linear_set_fused_policy(int(mode), int(max_m)), "vq_linear_set_fused_policy")


def linear_launch_count(G, rows, K):
    """1 when vq_linear_w8a8 runs this shape as the single fused kernel, 2 for quantise pass + GEMM."""
    return _lib.lib().vq_linear_launch_count(G, rows, K)


@_nvtx
def linear_w8a8(x, w: PreparedWeight, n_bits=8, smooth=None, ln=None, rows_per_mod=None, epi=VQ_EPI_BIAS, res=None,
                gate=None, rows_per_gate=0, out=None, ldo=None):
    """One QuantLayer-family forward in ONE call (vq_linear_w8a8): x fp16 [G, rows, K] (per-token statistics pooled over G)
    -> dynamic activation quantiser -> INT8 GEMM on the prepared weight -> dequant + epilogue -> fp16 [G*rows, N].
    ln = (shift, scale): LayerNorm + t2i_modulate in front (rows_per_mod as in ln_modulate_act_quant).  Shapes the fused
    kernel covers (K = 1152, small M) are a single launch; the rest is quantise pass + GEMM through a cached workspace."""
    _need_cuda_f16(x, "x")
    G, rows, K = x.shape
    if K != w.K:
        raise _lib.VqError(f"linear_w8a8: K mismatch {K} vs {w.K}")
    M = G * rows
    if out is None:
        out = torch.empty((M, w.N), dtype=torch.float16, device=x.device)
    shift = scale = None
    rpm = 0
    if ln is not None:
        shift, scale = ln
        _need_cuda_f16(shift, "shift")
        _need_cuda_f16(scale, "scale")
        rpm = rows if rows_per_mod is None else int(rows_per_mod)
        if shift.numel() != (M // rpm) * K or scale.numel() != shift.numel() or (rpm != rows and G != 1):
            raise _lib.VqError(f"linear_w8a8: shift/scale of {shift.numel()} elements do not match G={G} rows={rows} "
                               f"rows_per_mod={rpm} K={K}")
    if smooth is not None:
        _need_cuda_f16(smooth, "smooth")
    if epi == VQ_EPI_GATE_RESIDUAL:
        _need_cuda_f16(res, "res")
        _need_cuda_f16(gate, "gate")
    L = _lib.lib()
    n_launch = L.vq_linear_launch_count(G, rows, K)
    if n_launch == 1 and w.packed is not None and os.environ.get("VQ_W4_PACKED", "1") != "0":
        # W4A8: stream the weight operand as packed INT4 (half the bytes), expanded by the kernel's converter warps
        rc = L.vq_linear_w4a8(_ptr(x), G, rows, K, _ptr(smooth), _ptr(shift), _ptr(scale), rpm, n_bits, _ptr(w.packed),
                              _ptr(w.col), w.N, epi, _ptr(res), w.N, _ptr(gate), rows_per_gate, _ptr(out),
                              w.N if ldo is None else ldo, None, None, _ptr(status_word(x.device)), _stream())
        _lib.check(rc, "vq_linear_w4a8")
        _count()
        return out
    # scratch of the two-launch and of the overlapped schedule (codes, per-row parameters, panel flags); the panel-resident
    # fused kernel ignores it
    ws_bytes = L.vq_linear_workspace_bytes(G, rows, K)
    ws = _workspace(ws_bytes, x.device)
    rc = L.vq_linear_w8a8(_ptr(x), G, rows, K, _ptr(smooth), _ptr(shift), _ptr(scale), rpm, n_bits, _ptr(w.codes),
                          _ptr(w.col), w.N, epi, _ptr(res), w.N, _ptr(gate), rows_per_gate, _ptr(out),
                          w.N if ldo is None else ldo, None, None, _ptr(ws), ws_bytes, _ptr(status_word(x.device)),
                          _stream())
    _lib.check(rc, "vq_linear_w8a8")
    _count(n_launch)
    return out


@_nvtx
def attn_temporal(qkv, B, T, S, H, head_dim, scale, out=None):
    """qkv: fp16 [B*T*S, 3*H*head_dim] (fused q|k|v GEMM output, (T S) token order) -> fp16 [B*T*S, H*head_dim]."""
    _need_cuda_f16(qkv, "qkv")
    C = H * head_dim
    if qkv.shape != (B * T * S, 3 * C):
        raise _lib.VqError(f"attn_temporal: qkv shape {tuple(qkv.shape)} != {(B * T * S, 3 * C)}")
    if out is None:
        out = torch.empty((B * T * S, C), dtype=torch.float16, device=qkv.device)
    rc = _lib.lib().vq_attn_temporal(_ptr(qkv), _ptr(out), B, T, S, H, head_dim, float(scale), _stream())
    _lib.check(rc, "vq_attn_temporal")
    _count()
    return out


def attn_temporal_quant_supported(T, H, head_dim):
    return head_dim == 72 and H == 16 and 0 < T <= 16


@_nvtx
def attn_temporal_quant(qkv, B, T, S, H, head_dim, scale, n_bits=8, smooth=None) -> ActCodes:
    """attn_temporal + the per-token dynamic quantiser of the projection in one kernel: qkv fp16 [B*T*S, 3*H*head_dim] ->
    ActCodes of the attention output in (T S) token order (G = 1: every token on its own statistics)."""
    _need_cuda_f16(qkv, "qkv")
    C = H * head_dim
    if qkv.shape != (B * T * S, 3 * C):
        raise _lib.VqError(f"attn_temporal_quant: qkv shape {tuple(qkv.shape)} != {(B * T * S, 3 * C)}")
    if smooth is not None:
        _need_cuda_f16(smooth, "smooth")
    a = _alloc_act(1, B * T * S, C, qkv.device)
    rc = _lib.lib().vq_attn_temporal_quant(_ptr(qkv), B, T, S, H, head_dim, float(scale), _ptr(smooth), n_bits, _ptr(a.codes),
                                           _ptr(a.delta), _ptr(a.zp), _ptr(a.rowsum), _ptr(status_word(qkv.device)), _stream())
    _lib.check(rc, "vq_attn_temporal_quant")
    _count()
    return a


@_nvtx
def attn_spatial(qkv, n_seq, S, H, head_dim, scale, out=None):
    """qkv: fp16 [n_seq*S, 3*H*head_dim] (fused q|k|v GEMM output; n_seq = B*T frames of S tokens) -> fp16
    [n_seq*S, H*head_dim], token-major. tcgen05 flash attention; head_dim 72, S a multiple of 256."""
    _need_cuda_f16(qkv, "qkv")
    C = H * head_dim
    if qkv.shape != (n_seq * S, 3 * C):
        raise _lib.VqError(f"attn_spatial: qkv shape {tuple(qkv.shape)} != {(n_seq * S, 3 * C)}")
    if out is None:
        out = torch.empty((n_seq * S, C), dtype=torch.float16, device=qkv.device)
    rc = _lib.lib().vq_attn_spatial(_ptr(qkv), _ptr(out), n_seq, S, H, head_dim, float(scale), _stream())
    _lib.check(rc, "vq_attn_spatial")
    _count()
    return out


_i8_workspaces = {}    # (device index, stream handle) -> operand workspace of attn_spatial_i8 (separate from the linear scratch)


def attn_i8_workspace_layout(n_seq, S, H, head_dim=72):
    """Byte offsets of the pieces inside vq_attn_spatial_i8's workspace (mirrors ia_layout in vq_attn_i8.cu); tests read
    the codes and scales through it."""
    rows, C = n_seq * S, H * head_dim
    up = lambda v: (v + 255) & ~255
    off = {"qk8": 0}
    off["vt8"] = up(rows * 2 * H * 80)
    off["sq"] = up(off["vt8"] + n_seq * C * S)
    off["sk"] = up(off["sq"] + rows * H * 4)
    off["sv"] = up(off["sk"] + rows // 64 * H * 4)
    off["kmean"] = up(off["sv"] + n_seq * C * 4)
    off["svi"] = up(off["kmean"] + n_seq * C * 4)
    off["total"] = up(off["svi"] + n_seq * C * 4)
    return off


@_nvtx
def attn_spatial_i8(qkv, n_seq, S, H, head_dim, scale, out=None, workspace=None):
    """OPT-IN INT8 Q/K/V attention (vq_attn_spatial_i8; no reference counterpart, own tolerance — DESIGN.md 4.2d).
    Same contract as attn_spatial: qkv fp16 [n_seq*S, 3*H*72] -> fp16 [n_seq*S, H*72]; three launches (per-sequence
    statistics, operand codes, the tcgen05 kind::i8 attention kernel)."""
    _need_cuda_f16(qkv, "qkv")
    C = H * head_dim
    if qkv.shape != (n_seq * S, 3 * C):
        raise _lib.VqError(f"attn_spatial_i8: qkv shape {tuple(qkv.shape)} != {(n_seq * S, 3 * C)}")
    L = _lib.lib()
    nbytes = L.vq_attn_i8_workspace_bytes(n_seq, S, H, head_dim)
    if nbytes < 0:
        raise _lib.VqError(f"attn_spatial_i8: unsupported shape S={S} head_dim={head_dim}")
    if workspace is None:
        key = (qkv.device.index, _stream())
        workspace = _i8_workspaces.get(key)
        if workspace is None or workspace.numel() < nbytes:
            workspace = _i8_workspaces[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=qkv.device)
    elif workspace.numel() < nbytes:
        raise _lib.VqError(f"attn_spatial_i8: workspace of {workspace.numel()} bytes < {nbytes}")
    if out is None:
        out = torch.empty((n_seq * S, C), dtype=torch.float16, device=qkv.device)
    rc = L.vq_attn_spatial_i8(_ptr(qkv), _ptr(out), _ptr(workspace), n_seq, S, H, head_dim, float(scale), _stream())
    _lib.check(rc, "vq_attn_spatial_i8")
    _count(3)
    return out


def attn_spatial_supported(S, head_dim):
    return head_dim == 72 and S >= 256 and S % 256 == 0


@_nvtx
def attn_cross(q, kv, kv_start, kv_len, B, N, H, head_dim, max_len, scale, out=None):
    """q: fp16 [B*N, C]; kv: fp16 [sum(len), 2C]; kv_start/kv_len: int32 device tensors [B] -> fp16 [B*N, C]."""
    _need_cuda_f16(q, "q")
    _need_cuda_f16(kv, "kv")
    C = H * head_dim
    if out is None:
        out = torch.empty((B * N, C), dtype=torch.float16, device=q.device)
    rc = _lib.lib().vq_attn_cross(_ptr(q), _ptr(kv), _ptr(out), _ptr(kv_start), _ptr(kv_len), B, N, H, head_dim,
                                  int(max_len), int(kv.shape[0]), float(scale), _stream())
    _lib.check(rc, "vq_attn_cross")
    _count()
    return out
