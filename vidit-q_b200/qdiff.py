"""Host-side mirror of the reference's quantised-operator API (qdiff/), backed by the sm_100a kernels.

Same class names, constructor arguments, state attributes and checkpoint format as the reference so that scripts and
PTQ-calibrated `ckpt.pth` files work unchanged:

    QuantModel(model, weight_quant_params, act_quant_params, model_type)      qdiff/models/quant_model.py:38
    QuantLayer / QuantSpatialAttnLinear / QuantTemporalAttnLinear / QuantCrossAttnLinear          (STDiT)
    QuantAttnLinearImg / QuantCrossAttnLinearImg                                                   (PixArt)
    load_quant_params(qnn, ckpt_path)                                          qdiff/utils.py:65

What differs is what `forward` executes: when both weight and activation quantisation are enabled the layer runs
  vq_act_quant (per-token dynamic u8 codes) -> vq_gemm_w8a8 (tcgen05 INT8 GEMM + dequant epilogue)
on prepared u8 weight codes, instead of fake-quantising in fp16 and calling F.linear.  Layers left in floating point
(`remain_fp.txt`) run F.linear as in the reference.  PTQ-time work lives in viditq_b200.ptq (smooth-quant statistics,
weight parameters, static activation calibration); weight-only simulation and learned rounding are out of scope and raise.
Nothing here falls back to a CPU or fake-quant path.
"""
import logging
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

logger = logging.getLogger(__name__)


def find_interval(timerange, timestep_id):
    """Index of the [lo, hi] interval containing timestep_id (reference quant_layer.py:15)."""
    for i, (lo, hi) in enumerate(timerange):
        if lo <= timestep_id <= hi:
            return i
    return None


def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


# ---------------------------------------------------------------------------------------------------------------------
# quantiser state holders (buffers named exactly like the reference's so ckpt.pth round-trips)
# ---------------------------------------------------------------------------------------------------------------------
class BaseQuantizer(nn.Module):
    """State of one uniform-affine quantiser (reference base_quantizer.py:13-72). Holds parameters; the arithmetic
    lives in the CUDA kernels."""

    def __init__(self, quant_config):
        super().__init__()
        self.n_bits = _cfg_get(quant_config, "n_bits")
        self.mixed_precision = _cfg_get(quant_config, "mixed_precision")
        self.timestep_wise = _cfg_get(quant_config, "timestep_wise")
        self.bit_idx = self.mixed_precision.index(self.n_bits) if self.mixed_precision is not None else 0
        self.cur_timestep_id = 0
        self.per_group = _cfg_get(quant_config, "per_group")
        self.channel_dim = _cfg_get(quant_config, "channel_dim", 0)
        self.scale_method = _cfg_get(quant_config, "scale_method")
        self.round_mode = _cfg_get(quant_config, "round_mode")
        self.sym = _cfg_get(quant_config, "sym", False)
        self.running_stat = _cfg_get(quant_config, "running_stat", False)
        self.n_bitwidth = len(self.mixed_precision) if self.mixed_precision is not None else 1
        for name in ("delta_list", "zero_point_list", "delta", "zero_point", "alpha"):
            self.register_buffer(name, None)
        self.init_done = False

    def bitwidth_refactor(self, refactored_bit: int):
        """Reference base_quantizer.py:319-325: changes n_bits / bit_idx only — delta and zero_point are NOT
        re-selected once init_done (quirk Q7), so an '8-bit' layer of a 4-bit-calibrated ckpt stays on the 4-bit grid."""
        assert 2 <= refactored_bit <= 16, "bitwidth not supported"
        self.n_bits = refactored_bit
        if self.mixed_precision is not None:
            self.bit_idx = self.mixed_precision.index(self.n_bits)

    def forward(self, x):
        raise NotImplementedError("viditq_b200 quantisers hold state only; the fused kernels do the arithmetic")

    def extra_repr(self):
        return f"bit={self.n_bits}, per_group={self.per_group}, sym={self.sym}"


class WeightQuantizer(BaseQuantizer):
    pass


class ActQuantizer(BaseQuantizer):
    pass


class DynamicActQuantizer(ActQuantizer):
    pass


def _is_dynamic(aq):
    """DynamicActQuantizer of this module OR of the reference (accelerate() shares the reference's quantiser objects)."""
    return isinstance(aq, DynamicActQuantizer) or type(aq).__name__ == "DynamicActQuantizer"


class StraightThrough(nn.Module):
    def forward(self, x):
        return x


# ---------------------------------------------------------------------------------------------------------------------
# QuantLayer family
# ---------------------------------------------------------------------------------------------------------------------
class QuantLayer(nn.Module):
    """Drop-in for reference QuantLayer (quant_layer.py:22): forward(input) on [B, n_token, C] fp16."""

    def __init__(self, org_module: nn.Linear, weight_quant_params=None, act_quant_params=None,
                 disable_act_quant: bool = False, act_quant_mode: str = "qdiff"):
        super().__init__()
        if not isinstance(org_module, nn.Linear):
            raise NotImplementedError("viditq_b200 accelerates nn.Linear QuantLayers (all STDiT/PixArt quantised layers)")
        self.weight_quant_params = weight_quant_params
        self.act_quant_params = act_quant_params
        self.in_features = org_module.in_features
        self.out_features = org_module.out_features
        self.weight = org_module.weight
        self.org_weight = org_module.weight
        self.bias = org_module.bias
        self.org_bias = org_module.bias
        self.org_module = org_module
        self.weight_quant = False
        self.act_quant = False
        self.act_quant_mode = act_quant_mode
        self.disable_act_quant = disable_act_quant
        if weight_quant_params is not None:
            self.weight_quantizer = WeightQuantizer(weight_quant_params)
        if act_quant_params is not None:
            dyn = _cfg_get(act_quant_params, "dynamic", False)
            self.act_quantizer = DynamicActQuantizer(act_quant_params) if dyn else ActQuantizer(act_quant_params)
        self.split = 0
        self.calibrating = False     # set by viditq_b200.ptq while static activation quantisers are being calibrated
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False
        self.cur_timestep_id = 0
        sq = _cfg_get(act_quant_params, "smooth_quant", {}) or {}
        self.smooth_quant = bool(_cfg_get(sq, "enable", False))
        if self.smooth_quant:
            self.timerange = _cfg_get(sq, "timerange", [[0, 1000]])
            prev = -1
            for lo, hi in self.timerange:
                assert lo == prev + 1
                prev = hi
            assert prev == 1000
            self.timerange_num = len(self.timerange)
            self.act_quantizer.register_buffer("act_scale", None)
            self.channel_wise_scale_type = _cfg_get(sq, "channel_wise_scale_type", "dynamic")
            self.smooth_quant_momentum = _cfg_get(sq, "momentum", 0)
            self.smooth_quant_alpha = _cfg_get(sq, "alpha", None)
            self.smooth_quant_running_stat = False
        self._prepared = {}  # (n_bits, timerange_id) -> ops.PreparedWeight
        self._gen = 0        # bumped whenever cached weight codes become stale (model-level caches key on it)
        self._wmax_pow = {}  # alpha -> (weight data_ptr, max_n |W|^(1-alpha)) for the smooth-quant channel scale

    # -- state -----------------------------------------------------------------------------------------------------
    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.weight_quant = weight_quant
        self.act_quant = act_quant

    def get_quant_state(self):
        return self.weight_quant, self.act_quant

    def invalidate_prepared(self):
        """Drop cached u8 weight codes (call after changing weights or weight-quantiser buffers)."""
        self._prepared.clear()
        self._wmax_pow.clear()
        self._gen += 1

    def _apply(self, fn, *a, **k):  # .cuda()/.half() move the source tensors: prepared codes are stale
        self._prepared = {}
        self._wmax_pow = {}
        self._gen = getattr(self, "_gen", 0) + 1
        return super()._apply(fn, *a, **k)

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _timerange_id(self):
        if not hasattr(self, "timerange"):
            return 0
        return find_interval(self.timerange, self.cur_timestep_id)

    def _alpha(self, tr_id):
        alpha = self.smooth_quant_alpha
        if isinstance(alpha, (list, tuple)) or type(alpha).__name__ == "ListConfig":
            alpha = alpha[tr_id]
        return alpha

    def _weight_colmax_pow(self, alpha):
        """max_n |W[n, k]|^(1 - alpha) of quant_layer.py:116,137 — a function of the weight only: evaluated once (torch
        half ops, as the reference evaluates it every call)."""
        w = self.weight
        hit = self._wmax_pow.get(alpha)
        if hit is None or hit[0] != (w.data_ptr(), w.dtype):
            hit = ((w.data_ptr(), w.dtype), w.abs().max(dim=0)[0].pow(1 - alpha))
            self._wmax_pow[alpha] = hit
        return hit[1]

    def smooth_mode(self):
        """None | 'cached' (momentum scale from the checkpoint's act_scale: a constant per timerange) | 'dynamic'
        (quant_layer.py:115-116: from the live input) | 'running' (quant_layer.py:118-126 with
        smooth_quant_running_stat=True: the act_scale EMA advances on every call — quirk Q17, what
        t2i/scripts/quant_txt2img.py:297-300 leaves switched on for blocks.27.mlp.fc2 at inference)."""
        if not self.smooth_quant:
            return None
        kind = self.channel_wise_scale_type
        if kind == "dynamic":
            return "dynamic"
        if "momentum" in kind:
            return "running" if getattr(self, "smooth_quant_running_stat", False) else "cached"
        raise NotImplementedError(f"smooth-quant channel_wise_scale_type {kind!r}")

    def channel_wise_scale(self, tr_id):
        """quant_layer.py:137: act_scale[tr]^alpha / max_n |W|^(1-alpha), evaluated with torch half ops exactly as the
        reference does (a constant per timerange once the checkpoint is loaded; cached with the prepared weight)."""
        if "momentum" not in self.channel_wise_scale_type:
            raise NotImplementedError("channel_wise_scale(tr) is the checkpoint-constant (momentum) scale; the 'dynamic' "
                                      "type depends on the input: live_channel_scale(input)")
        alpha = self._alpha(tr_id)
        act_scale = self.act_quantizer.act_scale[tr_id]
        s = act_scale.pow(alpha) / self._weight_colmax_pow(alpha)
        return s.reshape(-1).contiguous()

    def _update_running_act_scale(self, cur_act_scale, tr_id):
        """quant_layer.py:119-126 / :146-153: EMA of the per-channel |x| maxima, in the tensor's own dtype.  The
        reference's `act_scale[tr].abs().mean() == 0` first-call test is a host sync; here both outcomes are computed on
        the device and selected there, so the update can live inside a captured CUDA graph."""
        aq = self.act_quantizer
        if aq.act_scale is None:
            aq.act_scale = torch.zeros([self.timerange_num, *cur_act_scale.shape]).to(cur_act_scale)
        old = aq.act_scale[tr_id]
        m = self.smooth_quant_momentum
        ema = old * m + cur_act_scale * (1 - m)
        aq.act_scale[tr_id] = torch.where(old.abs().mean() == 0, cur_act_scale, ema)

    @staticmethod
    def _col_absmax(input, gelu=False):
        """`input.abs().max(dim=-2)[0]` of quant_layer.py:116,119,147 for any rank (2-D timestep embeddings, 3-D token
        tensors, 4-D caption tensors): the per-channel maxima over the second-to-last dimension, from vq_col_absmax."""
        lead, n, K = input.shape[:-2], input.shape[-2], input.shape[-1]
        x3 = input.reshape(-1, n, K)
        return ops.col_absmax(x3 if x3.is_contiguous() else x3.contiguous(), gelu=gelu).reshape(*lead, K)

    def live_channel_scale(self, input, gelu=False):
        """Channel scale of the two input-dependent smooth-quant modes, [K] fp16.  The per-channel |x| maxima come from
        vq_col_absmax (one pass, exact); the [G, K] -> [K] arithmetic behind it is the reference's, op for op, on tiny
        tensors.  gelu=True: `input` is the pre-activation and the statistics are those of gelu_tanh(input)."""
        tr = self._timerange_id()
        alpha = self._alpha(tr)
        colmax = self._col_absmax(input, gelu=gelu)                     # input.abs().max(dim=-2)[0]  -> [G, K]
        if self.smooth_mode() == "dynamic":
            s = colmax.pow(alpha).mean(dim=0, keepdim=True) / self._weight_colmax_pow(alpha)
        else:
            self._update_running_act_scale(colmax.mean(dim=0, keepdim=True), tr)
            s = self.act_quantizer.act_scale[tr].pow(alpha) / self._weight_colmax_pow(alpha)
        return s.reshape(-1).contiguous()

    def _weight_src_key(self):
        wq = self.weight_quantizer
        d = wq.delta
        return (self.weight.data_ptr(), self.weight.dtype, None if d is None else (d.data_ptr(), d._version, d.dtype))

    def prepared_weight(self, live_smooth=None):
        """u8 weight codes + per-channel records of the current (n_bits, timerange).  Cached unless `live_smooth` (the
        input-dependent channel scale of this very call) is given: then the weight is re-quantised now, which is what
        the reference does on every call anyway (quant_layer.py:178-185)."""
        wq = self.weight_quantizer
        tr = self._timerange_id() if self.smooth_quant else 0
        key = (wq.n_bits, tr)
        src = self._weight_src_key()
        hit = None if live_smooth is not None else self._prepared.get(key)
        if hit is not None and hit[0] == src:
            return hit[1]
        if wq.delta is None or not wq.init_done:
            raise RuntimeError("weight quantiser has no parameters: load a PTQ ckpt (load_quant_params) or run "
                               "QuantModel.init_weight_quant_params() first")
        if wq.sym or wq.per_group != "channel" or wq.n_bits > 8:
            raise NotImplementedError("fused path supports asymmetric per-output-channel weights, <= 8 bits")
        if live_smooth is not None:
            smooth = live_smooth
        else:
            smooth = self.channel_wise_scale(tr) if self.smooth_mode() == "cached" else None
        pw = ops.prep_weight(self.weight.data, wq.delta, wq.zero_point, n_bits=wq.n_bits, smooth=smooth,
                             bias=None if self.bias is None else self.bias.data)
        pw.smooth = smooth
        if wq.n_bits <= 4 and pw.K == 1152:
            ops.pack_u4(pw)      # W4: the fused one-launch linear streams the weight as packed INT4 (vq_linear_w4a8)
        if live_smooth is None:
            if hit is not None:      # the source tensors changed under a cached entry (reload / .to()): stale everywhere
                self._gen += 1
            self._prepared[key] = (src, pw)
        return pw

    def _pool_view(self, input):
        """(G, rows) of the per-token statistics pool; subclasses restate the reference's reshape tricks."""
        return input.shape[0], input.shape[1]

    def _check_act_quantizer(self):
        aq = self.act_quantizer
        if aq.sym or aq.n_bits > 8:
            raise NotImplementedError("fused path supports asymmetric activations of <= 8 bits")
        if _is_dynamic(aq):
            if aq.per_group != "token":
                raise NotImplementedError("dynamic activation quantisation is per-token (the ViDiT-Q W8A8 / W4A8 configs)")
        elif aq.per_group not in (False, None, "token") or aq.delta is None or not (aq.init_done or self.calibrating):
            raise NotImplementedError("static activation quantisation needs calibrated per-tensor / per-token delta and "
                                      "zero_point from a PTQ checkpoint (set_quant_params_dict + set_quant_init_done)")

    def calibrate_static_act(self, input):
        """PTQ-time (viditq_b200.ptq): what the reference's static ActQuantizer does on every call while init_done is False
        (base_quantizer.py:112-127 -> init_quant_params :146-228): (delta, zero_point) re-derived from the live tensor —
        one pair per tensor (`per_group: False`, w8a8_naive.yaml) or per token of the layer's pooled view ('token') —,
        for every bit-width of `mixed_precision`, through the min / max EMA (momentum 0.95) when `running_stat` is set.
        Quirk kept: with mixed precision the reference calls init_quant_params once PER BIT-WIDTH, so the EMA advances
        n_bitwidth times per forward.  Arithmetic in the tensor's own dtype, op for op."""
        aq = self.act_quantizer
        G, rows = self._pool_view(input)
        x = input.reshape(G, rows, input.shape[-1])
        if aq.per_group == "token":
            xr = x.permute(1, 0, 2).reshape(rows, -1)
        elif not aq.per_group:
            xr = x.reshape(-1)
        else:
            raise NotImplementedError(f"static activation calibration with per_group={aq.per_group!r}")
        if aq.scale_method not in ("min_max", "max") or aq.sym:
            raise NotImplementedError("static activation calibration: asymmetric min_max only (the shipped configs)")
        bits = aq.mixed_precision if aq.mixed_precision is not None else [aq.n_bits]
        momentum = 0.95 if aq.running_stat else None          # base_quantizer.py:47
        for i, b in enumerate(bits):
            x_min = xr.min(dim=-1)[0]
            x_min = torch.where(x_min > 0, torch.zeros_like(x_min), x_min)
            x_max = xr.max(dim=-1)[0]
            x_max = torch.where(x_max < 0, torch.zeros_like(x_max), x_max)
            if momentum:
                if getattr(aq, "x_min", None) is None:
                    aq.x_min, aq.x_max = x_min, x_max
                else:
                    aq.x_min = aq.x_min * momentum + x_min * (1 - momentum)
                    aq.x_max = aq.x_max * momentum + x_max * (1 - momentum)
                    x_min, x_max = aq.x_min, aq.x_max
            delta = (x_max - x_min) / (2 ** b - 1)
            if delta.min() < 1e-6:
                delta = torch.full_like(delta, 1e-6)
            zp = torch.round(-x_min / delta)
            shape = [1, rows, 1] if aq.per_group == "token" else [1, 1, 1]
            delta, zp = delta.reshape(shape), zp.reshape(shape)
            if aq.delta_list is None:
                aq.delta_list = torch.full([len(bits), 1] + shape, -1.0, dtype=delta.dtype, device=delta.device)
                aq.zero_point_list = torch.full([len(bits), 1] + shape, -1.0, dtype=zp.dtype, device=zp.device)
            aq.delta_list[i, 0] = delta
            aq.zero_point_list[i, 0] = zp
        aq.delta = aq.delta_list[aq.bit_idx, 0]
        aq.zero_point = aq.zero_point_list[aq.bit_idx, 0]

    def _static_act_params(self):
        """(delta, zp) of a calibrated ActQuantizer as flat fp16 CUDA tensors: 1 element (per_group False,
        w8a8_naive.yaml) or n_token elements (static per-token, buffers of shape [1, n_token, 1])."""
        aq = self.act_quantizer
        dev = self.weight.device
        return (aq.delta.reshape(-1).to(dev, torch.float16).contiguous(),
                aq.zero_point.reshape(-1).to(dev, torch.float16).contiguous())

    def _weight_for(self, input, gelu=False, independent=False):
        """The prepared weight of THIS call: cached, or re-quantised under the input-dependent smooth-quant scale."""
        if self.smooth_mode() in ("dynamic", "running"):
            if independent:
                raise NotImplementedError("input-dependent smooth-quant scales with stacked independent calls (the EMA / "
                                          "batch mean would mix the calls)")
            return self.prepared_weight(live_smooth=self.live_channel_scale(input, gelu=gelu))
        return self.prepared_weight()

    def quantize_input(self, input, gelu=False, independent=False, ln=None, rows_per_mod=None):
        """Activation quantisation of a [*, n, C] fp16 tensor -> ops.ActCodes, with `.pw` = the prepared weight the codes
        belong to (the cached one, or — for the input-dependent smooth-quant modes — the one re-quantised for this call).
        gelu=True: `input` is the pre-activation of the preceding nn.GELU(approximate="tanh"); the activation is applied
        inside the quantise pass (fused schedules only — the module graph applies GELU itself).
        independent=True: the batch entries are separate forward calls stacked along the batch (cfg_split's cond / uncond
        halves at one prompt each): nothing is pooled, every row gets its own statistics."""
        self._check_act_quantizer()
        pw = self._weight_for(input, gelu, independent)
        smooth = getattr(pw, "smooth", None)
        if not _is_dynamic(self.act_quantizer):
            # static scales (base_quantizer.py:112-144 with init_done): nothing is computed from the live tensor.
            # gelu / ln = (shift, scale): the fused schedule's one-pass forms (the transform in front of the quantiser)
            delta, zp = self._static_act_params()
            x = input if input.is_contiguous() else input.contiguous()
            if delta.numel() > 1:   # static per-token: index = token position inside the layer's pooled view
                G, rows = self._pool_view(input)
                if rows != delta.numel():
                    raise NotImplementedError(f"static per-token scales for {delta.numel()} tokens, input has {rows}")
            a = ops.act_quant_static(x, delta, zp, n_bits=self.act_quantizer.n_bits, smooth=smooth, gelu=gelu, ln=ln,
                                     rows_per_mod=rows_per_mod)
        elif ln is not None:
            raise NotImplementedError("quantize_input(ln=...) is the static-scale form; dynamic layers use "
                                      "ops.ln_modulate_act_quant")
        else:
            G, rows = self._pool_view(input)
            if independent:
                G, rows = 1, G * rows
            x = input.reshape(G, rows, input.shape[-1])
            if not x.is_contiguous():
                x = x.contiguous()
            a = ops.act_quant(x, n_bits=self.act_quantizer.n_bits, smooth=smooth, gelu=gelu)
        a.pw = pw
        return a

    # -- forward ---------------------------------------------------------------------------------------------------
    def forward(self, input: torch.Tensor, scale: float = 1.0, split: int = 0, smooth_quant_enable: bool = False):
        if split != 0 or self.split != 0:
            raise NotImplementedError("split quantisation (UNet skip-concat) does not occur in STDiT/PixArt")
        act_q = self.act_quant and not self.disable_act_quant
        if (not self.smooth_quant and getattr(self, "smooth_quant_running_stat", False)
                and "momentum" in getattr(self, "channel_wise_scale_type", "")):
            # quant_layer.py:141-153: statistics collection without scaling (calibration-time; kept for API parity)
            self._update_running_act_scale(self._col_absmax(input).mean(dim=0, keepdim=True), self._timerange_id())
        if self.weight_quant and act_q:
            if _is_dynamic(self.act_quantizer):
                # the whole QuantLayer forward as ONE call (vq_linear_w8a8): a single fused kernel where the shape allows
                self._check_act_quantizer()
                pw = self._weight_for(input)
                G, rows = self._pool_view(input)
                x = input.reshape(G, rows, input.shape[-1])
                if not x.is_contiguous():
                    x = x.contiguous()
                out = ops.linear_w8a8(x, pw, n_bits=self.act_quantizer.n_bits, smooth=getattr(pw, "smooth", None))
            else:
                if self.calibrating and not self.act_quantizer.init_done:
                    self.calibrate_static_act(input)      # PTQ: parameters from this call's tensor, then quantise with them
                a = self.quantize_input(input)
                out = ops.gemm_w8a8(a, a.pw)
            return out.view(*input.shape[:-1], self.out_features)
        if not self.weight_quant and not act_q:
            if self.smooth_quant:
                raise NotImplementedError("FP forward of a smooth-quant layer (calibration-time) is out of scope")
            return F.linear(input, self.org_weight, self.org_bias)
        raise NotImplementedError("weight-only / activation-only simulated quantisation is a PTQ-time mode; the B200 "
                                  "path runs W+A quantised or full-precision layers")

    def extra_repr(self):
        return f"in={self.in_features}, out={self.out_features}, wq={self.weight_quant}, aq={self.act_quant}"


class QuantSpatialAttnLinear(QuantLayer):
    """stdit_quant_layer.py:10: input (B*T, S, C); statistics over the view (B, T*S, C)."""

    def _pool_view(self, input):
        T = self.act_quant_params["n_temporal_token"]
        S = self.act_quant_params["n_spatial_token"]
        assert input.shape[1] == S
        return input.shape[0] // T, T * S


class QuantTemporalAttnLinear(QuantLayer):
    """stdit_quant_layer.py:101: input (B*S, T, C); statistics over the view (B, S*T, C)."""

    def _pool_view(self, input):
        T = self.act_quant_params["n_temporal_token"]
        S = self.act_quant_params["n_spatial_token"]
        assert input.shape[1] == T
        return input.shape[0] // S, S * T


class QuantCrossAttnLinear(QuantLayer):
    """stdit_quant_layer.py:192: q_linear/proj see (B, T*S, C); kv_linear sees (1, sum(len), C) -> one token per row."""


class QuantAttnLinearImg(QuantLayer):
    """dit_quant_layer.py:9 (PixArt): plain (B, N, C), no smooth-quant path."""


class QuantCrossAttnLinearImg(QuantLayer):
    """dit_quant_layer.py:34 (PixArt)."""


class BaseQuantBlock(nn.Module):
    """Placeholder for the reference's diffusers-only quant blocks (unused for STDiT/PixArt)."""


# ---------------------------------------------------------------------------------------------------------------------
# QuantModel
# ---------------------------------------------------------------------------------------------------------------------
def pattern_in(text, pattern):
    """Dotted-name match with '*' wildcards and '[a-b]' integer ranges (semantics of quant_model.py:14-36)."""
    pats = pattern.split(".")
    toks = text.split(".")
    for start in range(len(toks)):
        ok = True
        for j, p in enumerate(pats):
            if p == "*":
                continue
            if start + j >= len(toks):
                raise IndexError("pattern runs past the module name")  # the reference indexes out of range here too
            t = toks[start + j]
            if "[" in p and "]" in p:
                lo, hi = p[1:-1].split("-")
                if t not in [str(v) for v in range(int(lo), int(hi) + 1)]:
                    ok = False
                    break
            elif t != p:
                ok = False
                break
        if ok:
            return True
    return False


def _safe_pattern_in(text, pattern):
    try:
        return pattern_in(text, pattern)
    except IndexError:
        return False


class QuantModel(nn.Module):
    def __init__(self, model: nn.Module, weight_quant_params=None, act_quant_params=None, model_type="opensora",
                 **kwargs):
        super().__init__()
        self.weight_quant = weight_quant_params is not None
        self.act_quant = act_quant_params is not None
        self.model_type = model_type
        self.timestep_wise = _cfg_get(act_quant_params, "timestep_wise", False)
        self.model = model
        self.in_channels = model.in_channels
        if hasattr(model, "image_size"):
            self.image_size = model.image_size
        self.quant_layer_refactor(self.model, weight_quant_params, act_quant_params)
        self.quant_params_dict = {}

    # replacement rules of quant_model.py:63-103
    def quant_layer_refactor(self, module, weight_quant_params, act_quant_params, prefix=""):
        for name, child in module.named_children():
            full = prefix + name if prefix else name
            if isinstance(child, nn.Linear):
                if ".attn." in full:
                    cls = QuantSpatialAttnLinear if self.model_type == "opensora" else QuantAttnLinearImg
                elif "cross_attn" in full:
                    cls = QuantCrossAttnLinear if self.model_type == "opensora" else QuantCrossAttnLinearImg
                elif "attn_temp" in full:
                    cls = QuantTemporalAttnLinear
                else:
                    cls = QuantLayer
                setattr(module, name, cls(child, weight_quant_params, act_quant_params))
            elif isinstance(child, (nn.Conv1d, nn.Conv2d)):
                if self.model_type == "opensora":
                    raise AssertionError("only linear layers are quantised in the STDiT model")
                # PixArt's x_embedder conv stays FP via the fp list in every shipped script; leave it untouched
            elif isinstance(child, (StraightThrough, QuantLayer)):
                continue
            else:
                self.quant_layer_refactor(child, weight_quant_params, act_quant_params, prefix=full + ".")

    def quant_layers(self):
        for name, m in self.model.named_modules():
            if isinstance(m, QuantLayer):
                yield name, m

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.weight_quant = weight_quant
        self.act_quant = act_quant
        for _, m in self.quant_layers():
            m.set_quant_state(weight_quant, act_quant)
        if hasattr(self, "fp_layer_list"):
            self.set_layer_quant(model=self, module_name_list=self.fp_layer_list, quant_level="per_layer",
                                 weight_quant=False, act_quant=False, prefix="")

    def get_quant_state(self):
        return self.weight_quant, self.act_quant

    def set_module_name_for_quantizer(self, module, prefix=""):
        for name, child in module.named_children():
            full = prefix + name if prefix else name
            if isinstance(child, BaseQuantizer):
                child.module_name = full
            else:
                self.set_module_name_for_quantizer(child, prefix=full + ".")

    def set_timestep_for_quantizer(self, t, module=None):
        for m in (self if module is None else module).modules():
            if isinstance(m, BaseQuantizer):
                m.cur_timestep_id = t

    def set_timestep_id_for_quantlayer(self, t, module=None):
        for m in (self if module is None else module).modules():
            if isinstance(m, QuantLayer):
                m.cur_timestep_id = t

    def set_quant_init_done(self, quantizer_type_name, module=None):
        kind = {"weight": WeightQuantizer, "activation": ActQuantizer}.get(quantizer_type_name)
        if kind is None:
            raise NotImplementedError
        for m in (self.model if module is None else module).modules():
            if isinstance(m, kind):
                m.init_done = True

    # ckpt format of quant_model.py:220-269: {quantizer module_name: [buffers OrderedDict, parameters OrderedDict]}
    def get_quant_params_dict(self, module=None, prefix="", dtype=torch.float32):
        if module is None:
            module = self.model
            self.quant_params_dict = {}
        for name, child in module.named_children():
            full = prefix + name if prefix else name
            if isinstance(child, BaseQuantizer):
                self.quant_params_dict[child.module_name] = [child._buffers, child._parameters]
            else:
                self.get_quant_params_dict(module=child, prefix=full + ".")
        return self.quant_params_dict

    def set_quant_params_dict(self, quant_params_dict, module=None, load_buffer_only=True, dtype=torch.float32):
        if module is None:
            module = self.model
        for _, child in module.named_children():
            if isinstance(child, BaseQuantizer):
                entry = quant_params_dict[child.module_name]
                if load_buffer_only:
                    assert len(entry[1]) == 0
                for bname, val in entry[0].items():
                    setattr(child, bname, val.to(dtype) if val is not None else None)
            else:
                self.set_quant_params_dict(quant_params_dict, module=child, load_buffer_only=load_buffer_only,
                                           dtype=dtype)
        for _, m in self.quant_layers():
            m.invalidate_prepared()

    def set_smooth_quant(self, smooth_quant, smooth_quant_running_stat):
        self.smooth_quant_stat = smooth_quant_running_stat
        for _, m in self.quant_layers():
            m.smooth_quant = smooth_quant
            m.smooth_quant_running_stat = smooth_quant_running_stat
            m.invalidate_prepared()

    def set_layer_smooth_quant(self, model, module_name_list, smooth_quant, smooth_quant_running_stat, prefix=""):
        for name, module in model.named_children():
            full = prefix + name if prefix else name
            if isinstance(module, QuantLayer):
                for pat in module_name_list:
                    if _safe_pattern_in(full, pat) or _safe_pattern_in(full, "model." + pat):
                        module.smooth_quant_running_stat = smooth_quant_running_stat
                        module.smooth_quant = smooth_quant
                        module.invalidate_prepared()
            else:
                self.set_layer_smooth_quant(module, module_name_list, smooth_quant, smooth_quant_running_stat,
                                            prefix=full + ".")

    def set_layer_quant(self, model=None, module_name_list=(), group_list=(), group_ignore=(), quant_level="per_layer",
                        weight_quant=True, act_quant=False, prefix=""):
        if quant_level not in ("per_layer", "per_group"):
            raise NotImplementedError("per_block quant levels address diffusers blocks (unused for STDiT/PixArt)")
        for name, module in model.named_children():
            full = prefix + name if prefix else name
            if isinstance(module, QuantLayer):
                if quant_level == "per_layer":
                    for pat in module_name_list:
                        if _safe_pattern_in(full, pat) or _safe_pattern_in(full, "model." + pat):
                            module.set_quant_state(weight_quant=weight_quant, act_quant=act_quant)
                else:
                    for cls_name in group_list:
                        hit = cls_name in full
                        if cls_name == "attn":
                            hit = hit and "cross_attn" not in full and "attn_temp" not in full
                        if hit and all(e not in full for e in group_ignore):
                            module.set_quant_state(weight_quant=weight_quant, act_quant=act_quant)
            else:
                self.set_layer_quant(model=module, module_name_list=module_name_list, group_list=group_list,
                                     group_ignore=group_ignore, quant_level=quant_level, weight_quant=weight_quant,
                                     act_quant=act_quant, prefix=full + ".")

    def load_bitwidth_config(self, model, bit_config, bit_type, prefix=""):
        """quant_model.py:562-586: per-layer bit-width table -> quantizer.bitwidth_refactor."""
        for name, module in model.named_children():
            full = prefix + name if prefix else name
            if isinstance(module, QuantLayer):
                if full in bit_config.keys():
                    if bit_type == "weight":
                        module.weight_quantizer.bitwidth_refactor(bit_config[full])
                    elif bit_type == "act":
                        module.act_quantizer.bitwidth_refactor(bit_config[full])
            else:
                self.load_bitwidth_config(model=module, bit_config=bit_config, bit_type=bit_type, prefix=full + ".")

    @torch.no_grad()
    def init_weight_quant_params(self, only_enabled=False, dtype=torch.float32):
        """Min-max per-output-channel weight parameters for every bit-width in `mixed_precision` — what the reference's
        PTQ weight pass (ptq.py:266-294 -> base_quantizer.py:166-228, 'channel' branch) stores into ckpt.pth.  Load-time
        torch ops on the weight's device; used by viditq_b200.ptq and so that synthetic-weight models can be benchmarked
        without a PTQ run.
        only_enabled: skip layers whose weight quantisation is off (the reference's calibration forward never initialises
        the remain_fp layers' quantisers: their buffers stay None in ckpt.pth).
        dtype: the arithmetic type.  The reference computes in the MODEL's dtype, op for op (x.min / x.max,
        (max - min) / (2^b - 1), round(-min / delta)) — fp16 when ptq.py runs with `dtype = "fp16"` (its 16x512x512
        config), fp32 otherwise; dtype=None reproduces that (the weight's own dtype).  The fp32 default is the
        higher-precision variant used for synthetic-weight benchmarks (pinned against an fp32 reference run)."""
        for _, m in self.quant_layers():
            if only_enabled and not m.weight_quant:
                continue
            wq = m.weight_quantizer
            wdt = m.weight.dtype if dtype is None else dtype
            w = m.weight.data.to(wdt)
            bits = wq.mixed_precision if wq.mixed_precision is not None else [wq.n_bits]
            n_t = len(m.timerange) if m.smooth_quant else 1
            dl = torch.empty(len(bits), n_t, w.shape[0], 1, device=w.device, dtype=wdt)
            zl = torch.empty_like(dl)
            for t in range(n_t):
                # quant_layer.py:178: weight_quantizer(self.weight * channel_wise_scale)
                wt = w * m.channel_wise_scale(t).to(wdt)[None, :] if m.smooth_quant else w
                mn = wt.min(dim=-1)[0]
                mn = torch.where(mn > 0, torch.zeros_like(mn), mn)
                mx = wt.max(dim=-1)[0]
                mx = torch.where(mx < 0, torch.zeros_like(mx), mx)
                for i, b in enumerate(bits):
                    delta = (mx - mn) / (2 ** b - 1)
                    if delta.min() < 1e-6:
                        delta = torch.full_like(delta, 1e-6)
                    dl[i, t, :, 0] = delta
                    zl[i, t, :, 0] = torch.round(-mn / delta)
            wq.delta_list, wq.zero_point_list = dl, zl
            wq.delta = dl[wq.bit_idx, 0].to(m.weight.dtype)
            wq.zero_point = zl[wq.bit_idx, 0].to(m.weight.dtype)
            wq.init_done = True
            m.invalidate_prepared()

    def forward(self, x, t, y, **kwargs):
        """quant_model.py:337-360: broadcast the scalar timestep to every QuantLayer, then run the wrapped model."""
        t0 = t[0].item() if isinstance(t, torch.Tensor) else float(t)
        if self.timestep_wise:
            self.set_timestep_for_quantizer(t0)
        self.set_timestep_id_for_quantlayer(t0)
        return self.model(x, t, y, **kwargs)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.model, name)


@torch.no_grad()
def load_quant_params(qnn, ckpt_path, dtype=torch.float32):
    """qdiff/utils.py:65-70."""
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.set_quant_params_dict(ckpt, dtype=dtype)


_LAYER_STATE = ("weight_quant", "act_quant", "disable_act_quant", "cur_timestep_id", "smooth_quant",
                "smooth_quant_running_stat", "smooth_quant_alpha", "smooth_quant_momentum", "channel_wise_scale_type",
                "timerange", "timerange_num", "split")


def accelerate(qnn):
    """Swap the forward of every *reference* QuantLayer inside `qnn` (an unmodified qdiff QuantModel) for the fused
    kernels, in place.  The reference module objects and all QuantModel methods stay as they are.

    The twin layer built here owns nothing: it wraps the same `org_module` (weights), holds the reference layer's OWN
    quantiser objects (so load_quant_params / set_quant_params_dict / load_bitwidth_config / set_quant_init_done keep
    acting on the state the kernels read, whether they run before or after accelerate()), and copies the layer-level
    switches the QuantModel API flips (set_quant_state, set_layer_quant, set_smooth_quant, set_layer_smooth_quant,
    timestep propagation) from the reference layer at the top of every call.  Prepared u8 weight codes are cached per
    (n_bits, timerange) and re-validated against the weight / delta tensors they were made from, so .cuda(), .half() and
    a checkpoint reload after accelerate() are picked up.  The running-stat smooth-quant EMA (quirk Q17) writes the
    reference quantiser's own `act_scale` buffer."""
    mapping = {"QuantLayer": QuantLayer, "QuantSpatialAttnLinear": QuantSpatialAttnLinear,
               "QuantTemporalAttnLinear": QuantTemporalAttnLinear, "QuantCrossAttnLinear": QuantCrossAttnLinear,
               "QuantAttnLinearImg": QuantAttnLinearImg, "QuantCrossAttnLinearImg": QuantCrossAttnLinearImg}
    n = 0
    for mod in qnn.model.modules():
        cls = mapping.get(type(mod).__name__)
        if cls is None or isinstance(mod, QuantLayer) or not hasattr(mod, "weight_quantizer"):
            continue
        if not isinstance(mod.org_module, nn.Linear):
            continue        # Conv QuantLayers (PixArt's x_embedder) stay on the reference path: FP list in every script
        ours = cls(mod.org_module, mod.weight_quant_params, mod.act_quant_params)
        # share the reference's quantiser objects (plain attribute entries: no second registration as sub-modules)
        for qname in ("weight_quantizer", "act_quantizer"):
            ours._modules.pop(qname, None)
            object.__setattr__(ours, qname, getattr(mod, qname))

        def fwd(input, scale=1.0, split=0, _ours=ours, _ref=mod, **kw):
            for name in _LAYER_STATE:
                if hasattr(_ref, name):
                    object.__setattr__(_ours, name, getattr(_ref, name))
            return _ours(input, scale, split)
        mod.forward = fwd
        object.__setattr__(mod, "_viditq_b200", ours)   # not a sub-module of the reference model (state_dict unchanged)
        n += 1
    return n
