"""GPU parity tests (run on the B200 with `pytest -m gpu`): the CUDA path, called through the C ABI, against
(a) vectors produced by the unmodified reference (tests/golden) and (b) the CPU oracle on seeded inputs.

Bar: integer codes / delta / zero-point / row sums bit-exact; GEMM output equal to the oracle's integer decomposition
(same operation order) and within 1e-3 relative (inf-norm and L2) of the reference's fake-quant fp16 output.
"""
import numpy as np
import pytest
import torch

from oracle import qdiff_oracle as O

pytestmark = pytest.mark.gpu

ACT_CASES = ["act/basic", "act/pooled_b2", "act/heavy_tail", "act/ragged_kv", "act/fc2_k4608",
             "act/signs_and_ranges", "act/bits6", "act/bits4"]
LAYER_VIEWS = {
    "layer/mlp_fc1": None, "layer/mlp_fc2_k4608": None, "layer/spatial_attn": 4, "layer/spatial_attn_b2": 4,
    "layer/temporal_attn": 16, "layer/cross_q": None, "layer/cross_kv": None, "layer/pixart_qkv_b2": None,
    "layer/pixart_cross_kv": None, "layer/nobias": None, "layer/w4_plain": None, "layer/w4_smooth_t100": 4,
    "layer/w4_smooth_t900": 4, "layer/w8_smooth_mlp_t700": None,
}


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200 import ops as _ops
    return _ops


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _pool_view(x, group):
    """(B*g, s, C) -> (B, g*s, C): the reshape the reference's spatial/temporal subclasses apply before quantising."""
    if group is None:
        return x
    return x.reshape(x.shape[0] // group, group * x.shape[1], x.shape[2])


def _smooth_for(c):
    if "act_scale" not in c:
        return None
    t = int(c["t_eval"])
    idx = next(i for i, (lo, hi) in enumerate(c["timerange"]) if lo <= t <= hi)
    return O.smooth_channel_scale(c["act_scale"][idx].reshape(-1), c["weight"], float(c["alpha"][idx]))


@pytest.mark.parametrize("name", ACT_CASES)
def test_act_quant_matches_reference_vectors(ops, golden, name):
    c = golden[name]
    x = dev(c["x"])
    a = ops.act_quant(x, n_bits=int(c["n_bits"]))
    torch.cuda.synchronize()
    B, n, C = c["x"].shape
    np.testing.assert_array_equal(a.delta.cpu().numpy(), c["delta"])
    np.testing.assert_array_equal(a.zp.cpu().numpy(), c["zp"])
    np.testing.assert_array_equal(a.codes.cpu().numpy().reshape(B, n, C), c["codes"])
    np.testing.assert_array_equal(a.rowsum.cpu().numpy().reshape(B, n), c["codes"].astype(np.int64).sum(-1))
    assert ops.check_status() == 0


@pytest.mark.parametrize("name", sorted(LAYER_VIEWS))
def test_quant_layer_family_matches_reference(ops, golden, name):
    c = golden[name]
    xv = _pool_view(c["x"], LAYER_VIEWS[name])
    B, n, C = xv.shape
    smooth = _smooth_for(c)
    w_bits = int(c["w_bits"])
    bias = c["bias"] if "bias" in c else None
    # weight prep (a2): codes and c1 bit-exact vs the oracle restatement of WeightQuantizer.forward
    pw = ops.prep_weight(dev(c["weight"]), dev(c["wdelta"]), dev(c["wzp"]), n_bits=w_bits,
                         smooth=None if smooth is None else dev(smooth), bias=None if bias is None else dev(bias))
    wq = O.weight_quant(c["weight"], c["wdelta"], c["wzp"], w_bits, smooth)
    np.testing.assert_array_equal(pw.codes.cpu().numpy(), wq["codes"])
    zw = np.rint(c["wzp"].astype(np.float32)).astype(np.int64)
    col = pw.col.cpu().numpy()
    np.testing.assert_array_equal(col[:, 0], wq["colsum"].astype(np.int64) - C * zw)
    np.testing.assert_array_equal(col[:, 1], zw)
    # act quant (a1): delta / zp as the reference left them in its act_quantizer buffers
    a = ops.act_quant(dev(xv), n_bits=8, smooth=None if smooth is None else dev(smooth))
    np.testing.assert_array_equal(a.delta.cpu().numpy(), c["adelta"])
    np.testing.assert_array_equal(a.zp.cpu().numpy(), c["azp"])
    oa = O.dynamic_act_quant(xv, 8, smooth)
    np.testing.assert_array_equal(a.codes.cpu().numpy().reshape(B, n, C), oa["codes"])
    # GEMM + dequant epilogue (a3-a7)
    y = ops.gemm_w8a8(a, pw).cpu().numpy().reshape(B, n, -1)
    yo = O.quant_linear_int(oa["codes"], oa["delta"], oa["zp"], oa["rowsum"], wq["codes"], c["wdelta"], c["wzp"], bias)
    mism = (y.view(np.uint16) != yo.view(np.uint16))
    assert mism.mean() <= 1e-5, mism.mean()      # fma vs float64 double rounding can flip a last bit, nothing else
    ref = c["out"].reshape(B, n, -1).astype(np.float32)
    yf = y.astype(np.float32)
    assert np.abs(yf - ref).max() / np.abs(ref).max() <= 1e-3
    assert np.linalg.norm(yf - ref) / np.linalg.norm(ref) <= 1e-3


def _rand_layer(seed, M, K, N, heavy=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((1, M, K)).astype(np.float32)
    if heavy:
        x[..., [3, 77, 500, 901]] *= 20.0
    w = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
    b = (rng.standard_normal(N) * 0.01).astype(np.float16)
    wd, wz = O.weight_init_params(w, 8)
    wd = (wd * rng.uniform(0.5, 1.5, size=N)).astype(np.float16)   # "random calib scales" (BASELINE config 1)
    return x.astype(np.float16), w.astype(np.float16), b, wd, wz.astype(np.float16)


@pytest.mark.parametrize("epi", ["bias", "gelu_tanh", "gate_residual"])
def test_gemm_epilogues_vs_oracle(ops, epi):
    M, K, N = 512, 1152, 384
    x, w, b, wd, wz = _rand_layer(1, M, K, N)
    a = ops.act_quant(dev(x))
    pw = ops.prep_weight(dev(w), dev(wd), dev(wz), bias=dev(b))
    oa = O.dynamic_act_quant(x)
    wq = O.weight_quant(w, wd, wz)
    rng = np.random.default_rng(5)
    res = rng.standard_normal((1, M, N)).astype(np.float16)
    gate = (rng.standard_normal((1, N)) * 0.5).astype(np.float16)
    code = {"bias": ops.VQ_EPI_BIAS, "gelu_tanh": ops.VQ_EPI_GELU_TANH, "gate_residual": ops.VQ_EPI_GATE_RESIDUAL}[epi]
    y = ops.gemm_w8a8(a, pw, epi=code, res=dev(res.reshape(M, N)), gate=dev(gate), rows_per_gate=M)
    yo = O.quant_linear_int(oa["codes"], oa["delta"], oa["zp"], oa["rowsum"], wq["codes"], wd, wz, b, epi=epi,
                            res16=res, gate16=gate)
    y = y.cpu().numpy().reshape(1, M, N)
    mism = (y.view(np.uint16) != yo.view(np.uint16))
    # identical integer accumulators; the only freedom is the last fp16 bit of y (fma vs float64 double rounding,
    # __expf vs float64 tanh in GELU), which the gated residual can carry through a cancellation
    assert mism.mean() <= (2e-3 if epi == "gelu_tanh" else 1e-5), mism.mean()
    scale = np.float16(np.abs(yo.astype(np.float32)).max() if epi != "gate_residual" else
                       max(np.abs(res.astype(np.float32)).max(), np.abs(yo.astype(np.float32)).max()))
    assert np.abs(y.astype(np.float32) - yo.astype(np.float32)).max() <= 2 * float(np.spacing(scale))


def test_config1_full_size_w8a8_linear(ops):
    """BASELINE config 1 at full size: QuantLinear W8A8 in=1152 out=4608, M=16384 tokens."""
    M, K, N = 16384, 1152, 4608
    x, w, b, wd, wz = _rand_layer(7, M, K, N)
    xd = dev(x)
    a = ops.act_quant(xd)
    pw = ops.prep_weight(dev(w), dev(wd), dev(wz), bias=dev(b))
    y = ops.gemm_w8a8(a, pw)
    torch.cuda.synchronize()
    oa = O.dynamic_act_quant(x)
    np.testing.assert_array_equal(a.codes.cpu().numpy().reshape(1, M, K), oa["codes"])
    np.testing.assert_array_equal(a.delta.cpu().numpy().astype(np.float32), oa["delta"])
    np.testing.assert_array_equal(a.zp.cpu().numpy().astype(np.float32), oa["zp"])
    np.testing.assert_array_equal(a.rowsum.cpu().numpy().reshape(1, M), oa["rowsum"])
    wq = O.weight_quant(w, wd, wz)
    np.testing.assert_array_equal(pw.codes.cpu().numpy(), wq["codes"])
    rows = np.random.default_rng(0).choice(M, 256, replace=False)
    rows.sort()
    sub = dict(codes=oa["codes"][:, rows], delta=oa["delta"][rows], zp=oa["zp"][rows], rowsum=oa["rowsum"][:, rows])
    yo = O.quant_linear_int(sub["codes"], sub["delta"], sub["zp"], sub["rowsum"], wq["codes"], wd, wz, b)
    ys = y.cpu().numpy()[rows].reshape(1, len(rows), N)
    assert (ys.view(np.uint16) != yo.view(np.uint16)).mean() <= 1e-5
    # against the reference-style fake-quant path (fp16 operands, fp32 accumulate) on the same rows
    fake = O.linear_f16(oa["xhat"][0, rows], wq["what"], b).astype(np.float32)
    yf = ys[0].astype(np.float32)
    assert np.abs(yf - fake).max() / np.abs(fake).max() <= 1e-3
    assert np.linalg.norm(yf - fake) / np.linalg.norm(fake) <= 1e-3
    # size-independent property: the integer form is exactly linear in the bias (shift by a constant vector)
    pw2 = ops.prep_weight(dev(w), dev(wd), dev(wz), bias=None)
    y0 = ops.gemm_w8a8(a, pw2).float()
    d = (y.float() - y0 - dev(b).float()[None, :]).abs().max().item()
    assert d <= 2 * float(np.spacing(np.float16(np.abs(yf).max())))


def test_pooled_batch_statistics_and_rows_period(ops):
    """Quirk Q1: PixArt-style batch-2 forward shares one (delta, zp) per token across the batch."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 256, 1152)).astype(np.float16)
    x[1] *= 4
    w, b = (rng.standard_normal((192, 1152)) * 0.02).astype(np.float16), None
    wd, wz = O.weight_init_params(w.astype(np.float32), 8)
    a = ops.act_quant(dev(x))
    oa = O.dynamic_act_quant(x)
    np.testing.assert_array_equal(a.codes.cpu().numpy().reshape(2, 256, 1152), oa["codes"])
    pw = ops.prep_weight(dev(w), dev(wd.astype(np.float16)), dev(wz.astype(np.float16)))
    y = ops.gemm_w8a8(a, pw).cpu().numpy().reshape(2, 256, 192)
    wq = O.weight_quant(w, wd.astype(np.float16), wz.astype(np.float16))
    yo = O.quant_linear_int(oa["codes"], oa["delta"], oa["zp"], oa["rowsum"], wq["codes"], wd.astype(np.float16),
                            wz.astype(np.float16), b)
    assert (y.view(np.uint16) != yo.view(np.uint16)).mean() <= 1e-5


@pytest.mark.parametrize("G,rows,K", [(1, 300, 1152), (2, 64, 1152), (1, 40, 4608)])
def test_ln_modulate_act_quant(ops, G, rows, K):
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((G, rows, K)) * 2 + 0.3).astype(np.float16)
    shift = (rng.standard_normal((G, K)) * 0.2).astype(np.float16)
    scale = (rng.standard_normal((G, K)) * 0.2).astype(np.float16)
    a, y = ops.ln_modulate_act_quant(dev(x), dev(shift), dev(scale), want_y=True)
    y = y.cpu().numpy()
    yo = O.ln_modulate(x, shift, scale)
    # fp32 vs fp64 LayerNorm statistics move h(LN) by at most one fp16 ulp; |LN * (1+scale)| <~ 8 bounds that ulp, and
    # the final add can cancel, so compare absolutely rather than in ulps of the (possibly tiny) result
    diff = np.abs(y.astype(np.float32) - yo.astype(np.float32))
    assert diff.max() <= 2 * 2.0 ** -8 and (diff > 0).mean() < 2e-2, (diff.max(), (diff > 0).mean())
    # given the kernel's own modulated tensor, the quantiser part must be bit-exact
    oa = O.dynamic_act_quant(y)
    np.testing.assert_array_equal(a.codes.cpu().numpy().reshape(G, rows, K), oa["codes"])
    np.testing.assert_array_equal(a.delta.cpu().numpy().astype(np.float32), oa["delta"])
    np.testing.assert_array_equal(a.zp.cpu().numpy().astype(np.float32), oa["zp"])
    np.testing.assert_array_equal(a.rowsum.cpu().numpy().reshape(G, rows), oa["rowsum"])


def test_eps_degenerate_row_sets_status_and_raises(ops):
    x = np.random.default_rng(0).standard_normal((1, 64, 1152)).astype(np.float16)
    x[0, 5] = 0
    ops.status_word().zero_()
    ops.act_quant(dev(x))
    with pytest.raises(Exception, match="delta < 1e-6"):
        ops.check_status()
    assert ops.check_status() == 0   # sticky bit cleared after being reported


def test_cpu_tensor_is_rejected_no_fallback(ops):
    with pytest.raises(Exception, match="no CPU path"):
        ops.act_quant(torch.zeros(1, 8, 64, dtype=torch.float16))


# ---------------------------------------------------------------------------------------------------- attention (fp16)
def _sdpa_ref(q, k, v, scale):
    """Plain fp32 softmax attention: q [n, Lq, H, D], k/v [n, Lk, H, D]."""
    qf, kf, vf = (t.float().transpose(1, 2) for t in (q, k, v))
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    return (p @ vf).transpose(1, 2)


@pytest.mark.parametrize("B,T,S", [(1, 16, 64), (2, 4, 33), (1, 7, 40)])
def test_temporal_attention_matches_fp32_reference(ops, B, T, S):
    H, D = 16, 72
    C = H * D
    torch.manual_seed(0)
    qkv = (torch.randn(B * T * S, 3 * C, device="cuda") * 1.5).half()
    out = ops.attn_temporal(qkv, B, T, S, H, D, D ** -0.5)
    v5 = qkv.view(B, T, S, 3, H, D)
    q, k, v = (v5[:, :, :, j].permute(0, 2, 1, 3, 4).reshape(B * S, T, H, D) for j in range(3))
    ref = _sdpa_ref(q, k, v, D ** -0.5).reshape(B, S, T, C).permute(0, 2, 1, 3).reshape(B * T * S, C)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-3 * ref.abs().max().item() + 1e-3, err      # fp16 P and fp16 output rounding


@pytest.mark.parametrize("n_seq,S,outliers", [(2, 1024, False), (3, 256, True), (5, 512, True), (33, 1024, False)])
def test_spatial_attention_matches_fp32_reference(ops, n_seq, S, outliers):
    """tcgen05 flash attention (vq_attn_spatial) vs plain fp32 softmax attention and vs the library flash kernel the
    reference calls (fp16 in / out): same tolerance as the library's own distance to fp32. `outliers` makes late keys
    larger so that the running maximum rises across key tiles (lazy-rescale path); 33 sequences make persistent CTAs walk
    several work items each."""
    H, D = 16, 72
    C = H * D
    torch.manual_seed(3)
    qkv = (torch.randn(n_seq * S, 3 * C, device="cuda") * 1.3)
    if outliers:
        v5 = qkv.view(n_seq, S, 3, H, D)
        v5[:, S // 2::7, 1] *= 3.0          # every 7th key of the second half
        v5[:, -1, 1, 3] *= 8.0              # one dominant last key for head 3
    qkv = qkv.half()
    out = ops.attn_spatial(qkv, n_seq, S, H, D, D ** -0.5)
    v5 = qkv.view(n_seq, S, 3, H, D)
    ref = _sdpa_ref(v5[:, :, 0], v5[:, :, 1], v5[:, :, 2], D ** -0.5).reshape(n_seq * S, C)
    lib = torch.nn.functional.scaled_dot_product_attention(
        v5[:, :, 0].transpose(1, 2), v5[:, :, 1].transpose(1, 2), v5[:, :, 2].transpose(1, 2), scale=D ** -0.5)
    lib = lib.transpose(1, 2).reshape(n_seq * S, C).float()
    scale_ref = ref.abs().max().item()
    err = (out.float() - ref).abs().max().item()
    err_lib = (lib - ref).abs().max().item()
    assert err <= 4e-3 * scale_ref + 1e-3, (err, err_lib)      # fp16 P and fp16 output rounding
    assert err <= 4 * err_lib + 1e-3, (err, err_lib)           # no worse than a few times the library's own error
    rel_l2 = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel_l2 <= 1e-3, rel_l2


def test_spatial_attention_rejects_unsupported_shapes(ops):
    from viditq_b200 import _lib
    qkv = torch.zeros(100, 3 * 1152, device="cuda", dtype=torch.float16)
    with pytest.raises(_lib.VqError):
        ops.attn_spatial(qkv, 1, 100, 16, 72, 72 ** -0.5)      # S not a multiple of 256: no silent fallback


@pytest.mark.parametrize("B,N,lens", [(1, 256, [109]), (2, 200, [77, 120]), (1, 130, [1]), (3, 48, [128, 5, 64]),
                                      (3, 512, [128, 5, 64]), (2, 1024, [1, 120]), (5, 768, [64, 65, 63, 2, 127])])
def test_cross_attention_matches_fp32_reference(ops, B, N, lens):
    H, D = 16, 72
    C = H * D
    torch.manual_seed(1)
    q = torch.randn(B * N, C, device="cuda").half()
    kv = torch.randn(sum(lens), 2 * C, device="cuda").half()
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    out = ops.attn_cross(q, kv, torch.from_numpy(starts).cuda(), torch.tensor(lens, dtype=torch.int32, device="cuda"),
                         B, N, H, D, max(lens), D ** -0.5)
    refs, off = [], 0
    for b in range(B):
        kk = kv[off:off + lens[b]].view(1, lens[b], 2, H, D)
        refs.append(_sdpa_ref(q[b * N:(b + 1) * N].view(1, N, H, D), kk[:, :, 0], kk[:, :, 1], D ** -0.5).reshape(N, C))
        off += lens[b]
    ref = torch.cat(refs, 0)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-3 * ref.abs().max().item() + 1e-3, err


@pytest.mark.parametrize("G,rows,K,smooth", [(1, 257, 4608, False), (2, 64, 4608, True), (1, 100, 1152, False)])
def test_gelu_act_quant_equals_gelu_then_quant(ops, G, rows, K, smooth):
    """GELU(tanh) fused in front of fc2's quantiser == the GELU epilogue's fp16 output quantised by the plain pass, and
    == the oracle's gelu -> fp16 -> (/ smooth) -> DynamicActQuantizer up to the last-bit freedom of the activation."""
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((G, rows, K)) * 2.0).astype(np.float16)
    x[..., [5, 300]] *= 6.0
    sm = rng.uniform(0.5, 2.0, size=K).astype(np.float16) if smooth else None
    a = ops.act_quant(dev(x), smooth=dev(sm) if smooth else None, gelu=True)
    g16 = O.gelu_tanh(x.astype(np.float64)).astype(np.float32).astype(np.float16)
    b = ops.act_quant(dev(g16), smooth=dev(sm) if smooth else None)
    # MUFU ex2/rcp vs float64 tanh: the fp16 activation may differ in its last bit on a few elements.  Where that
    # element is the row extremum the step size moves by one fp16 ulp (rare); elsewhere a code moves by +-1 (rare).
    same = ((a.delta == b.delta) & (a.zp == b.zp)).cpu().numpy()
    assert same.mean() >= 0.97, same.mean()
    ca = a.codes.cpu().numpy().astype(np.int16).reshape(G, rows, K)[:, same]
    cb = b.codes.cpu().numpy().astype(np.int16).reshape(G, rows, K)[:, same]
    assert np.abs(ca - cb).max() <= 1 and (ca != cb).mean() <= 2e-3, ((ca != cb).mean(), np.abs(ca - cb).max())
    dd = np.abs(a.delta.cpu().numpy().astype(np.float32) - b.delta.cpu().numpy().astype(np.float32))
    assert dd.max() <= float(np.spacing(np.float16(b.delta.float().max().item())))
    np.testing.assert_array_equal(a.rowsum.cpu().numpy(),
                                  a.codes.cpu().numpy().astype(np.int64).reshape(G * rows, K).sum(-1))


@pytest.mark.parametrize("G,T,S", [(1, 4, 64), (2, 3, 40)])
def test_act_quant_heads_equals_token_major(ops, G, T, S):
    """Head-major input path (attention output [G*T, H, S, 72]) == quantising the transposed token-major copy."""
    H, D = 16, 72
    torch.manual_seed(2)
    o = (torch.randn(G * T, H, S, D, device="cuda") * 1.3).half()
    tok = o.transpose(1, 2).reshape(G, T * S, H * D).contiguous()
    a = ops.act_quant_heads(o, G, T * S, S)
    b = ops.act_quant(tok)
    assert torch.equal(a.codes, b.codes) and torch.equal(a.delta, b.delta)
    assert torch.equal(a.zp, b.zp) and torch.equal(a.rowsum, b.rowsum)
    oa = O.dynamic_act_quant(tok.cpu().numpy())
    np.testing.assert_array_equal(a.codes.cpu().numpy().reshape(G, T * S, H * D), oa["codes"])


@pytest.mark.parametrize("n,k", [(1, 0.0), (2, 0.25)])
def test_cfg_ddim_step_is_bit_identical_to_the_eager_sequence(ops, n, k):
    """vq_cfg_ddim_step == forward_with_cfg's combine + ddim_sample(eta=0) as separate fp32 ops (sampler.py restatement,
    itself pinned to the reference by tests/golden/sampler_golden.npz)."""
    from viditq_b200.sampler import SpacedDDIM
    torch.manual_seed(4)
    ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)
    x = torch.randn(n, 4, 5, 16, 16, device="cuda")
    oc = torch.randn(n, 8, 5, 16, 16, device="cuda")
    ou = torch.randn(n, 8, 5, 16, 16, device="cuda")
    for i in (99, 37, 0):
        coef = ddim.coefficients(i, "cuda")
        want = SpacedDDIM.ddim_update(x, SpacedDDIM.cfg_combine(oc, ou, ddim.cfg_scale, ptqd_k=k), coef)
        got = ops.cfg_ddim_step(oc, ou, x, coef, ddim.cfg_scale, ptqd_k=k)
        assert torch.equal(got, want), (i, (got - want).abs().max().item())
    with pytest.raises(Exception):
        ops.cfg_ddim_step(oc.cpu(), ou.cpu(), x.cpu(), coef.cpu(), 4.0)


@pytest.mark.parametrize("name", ["static/tensor_mlp", "static/tensor_spatial_b2", "static/tensor_bits6",
                                  "static/token_mlp"])
def test_static_act_scales_match_reference(ops, golden_static, name):
    """vq_act_quant_static (N3, w8a8_naive.yaml): codes bit-exact against the unmodified reference's static
    ActQuantizer, and the layer output through vq_gemm_w8a8 (a_rows_period = number of (delta, zp) pairs) <= 1e-3."""
    c = golden_static[name]
    bits = int(c["a_bits"])
    x = c["x"]
    a = ops.act_quant_static(dev(x), dev(c["adelta"]), dev(c["azp"]), n_bits=bits)
    np.testing.assert_array_equal(a.codes.cpu().numpy().reshape(c["codes"].shape), c["codes"])
    np.testing.assert_array_equal(a.rowsum.cpu().numpy().reshape(c["codes"].shape[:-1]),
                                  c["codes"].astype(np.int64).sum(-1))
    pw = ops.prep_weight(dev(c["weight"]), dev(c["wdelta"]), dev(c["wzp"]), bias=dev(c["bias"]))
    y = ops.gemm_w8a8(a, pw).cpu().numpy().astype(np.float32).reshape(c["out"].shape)
    ref = c["out"].astype(np.float32)
    assert np.abs(y - ref).max() / np.abs(ref).max() <= 1e-3
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) <= 1e-3


def test_static_quant_layer_module_path(ops, golden_static):
    """The same through viditq_b200.qdiff.QuantLayer configured like w8a8_naive.yaml (per_group False, dynamic False) with
    the reference's calibrated buffers loaded the way set_quant_params_dict does."""
    import torch.nn as nn
    from oracle import ref_shims
    from viditq_b200 import qdiff
    c = golden_static["static/tensor_mlp"]
    N, K = c["weight"].shape
    lin = nn.Linear(K, N)
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(c["weight"].astype(np.float32)))
        lin.bias.copy_(torch.from_numpy(c["bias"].astype(np.float32)))
    wq, aq = ref_shims.w8a8_dynamic_configs(n_temporal=4, n_spatial=16, n_prompt=8)
    aq["dynamic"], aq["per_group"] = False, False
    layer = qdiff.QuantLayer(lin, wq, aq)
    layer.weight_quantizer.delta = torch.from_numpy(c["wdelta"].astype(np.float32)).reshape(-1, 1)
    layer.weight_quantizer.zero_point = torch.from_numpy(c["wzp"].astype(np.float32)).reshape(-1, 1)
    layer.act_quantizer.delta = torch.from_numpy(c["adelta"].astype(np.float32)).reshape(1, 1, 1)
    layer.act_quantizer.zero_point = torch.from_numpy(c["azp"].astype(np.float32)).reshape(1, 1, 1)
    layer.weight_quantizer.init_done = layer.act_quantizer.init_done = True
    layer.set_quant_state(True, True)
    layer.cuda().half()
    with torch.no_grad():
        y = layer(dev(c["x"])).cpu().numpy().astype(np.float32)
    ref = c["out"].astype(np.float32)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) <= 1e-3
    assert np.abs(y - ref).max() / np.abs(ref).max() <= 1e-3


# ---------------------------------------------------------------------------------------------------- patch embedding
@pytest.mark.parametrize("B,T,H,W", [(2, 16, 64, 64), (1, 3, 6, 10), (3, 1, 8, 8)])
def test_patch_embed_matches_conv3d_plus_pos(ops, B, T, H, W):
    """vq_patch_embed == Conv3d(kernel = stride = (1,2,2)) on the fp16-rounded latent + transpose + pos_embed add
    (stdit.py:255-258), up to the summation order inside the 16-term dot product (last fp16 bit)."""
    C, Cin = 1152, 4
    torch.manual_seed(5)
    conv = torch.nn.Conv3d(Cin, C, kernel_size=(1, 2, 2), stride=(1, 2, 2)).cuda().half()
    S = (H // 2) * (W // 2)
    pos = (torch.randn(S, C, device="cuda") * 0.5).half()
    z = torch.randn(B, Cin, T, H, W, device="cuda")
    out = ops.patch_embed(z, conv.weight, conv.bias, pos, (2, 2))
    # reference in fp32 on the fp16-rounded operands, rounded to fp16 where the fp16 graph rounds
    y = torch.nn.functional.conv3d(z.half().float(), conv.weight.float(), conv.bias.float(), stride=(1, 2, 2))
    y = y.half().flatten(2).transpose(1, 2).reshape(B, T, S, C)
    ref = (y + pos).reshape(B, T * S, C)           # fp16 add
    assert out.shape == ref.shape
    diff = (out.float() - ref.float()).abs()
    # one fp16 ulp of the larger of (conv output, position embedding, sum): the add can cancel
    mag = torch.maximum(torch.maximum(y.float().abs(), pos.float().abs().expand_as(y)).reshape(B, T * S, C), ref.float().abs())
    ulp = torch.clamp(mag, min=1e-3) * 2.0 ** -10
    assert bool((diff <= 2.02 * ulp).all()), float((diff / ulp).max())   # 1 ulp of the conv output, then 1 of the sum
    assert float((out != ref).float().mean()) < 2e-2     # and almost always bit-identical


@pytest.mark.parametrize("G,T,S,smooth", [(1, 16, 40, False), (2, 4, 33, True)])
def test_add_act_quant_equals_add_then_quant(ops, G, T, S, smooth):
    """vq_add_act_quant == the fp16 add of the temporal position embedding followed by the plain quantiser, bit for bit."""
    K = 1152
    torch.manual_seed(7)
    x = (torch.randn(G, T * S, K, device="cuda") * 2).half()
    tpe = torch.randn(T, K, device="cuda").half()
    sm = (torch.rand(K, device="cuda") + 0.5).half() if smooth else None
    a = ops.add_act_quant(x, tpe, S, smooth=sm)
    xt = (x.view(G, T, S, K) + tpe.view(1, T, 1, K)).view(G, T * S, K).contiguous()
    b = ops.act_quant(xt, smooth=sm)
    for u, v in ((a.codes, b.codes), (a.delta, b.delta), (a.zp, b.zp), (a.rowsum, b.rowsum)):
        assert torch.equal(u, v)


@pytest.mark.parametrize("B,T,S", [(1, 16, 256), (2, 16, 64), (1, 4, 64), (2, 7, 33)])
def test_temporal_attention_with_fused_quantiser_is_bit_identical(B, T, S):
    """vq_attn_temporal_quant (one block = all 16 heads of a position; the projection's per-token quantiser fused behind
    the attention) against vq_attn_temporal -> vq_act_quant: same codes, delta, zero points and row sums."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200 import ops
    H, D = 16, 72
    g = torch.Generator().manual_seed(B * 100 + T + S)
    qkv = (torch.randn(B * T * S, 3 * H * D, generator=g) * 1.3).half().cuda()
    qkv[:, 5] *= 6
    o = ops.attn_temporal(qkv, B, T, S, H, D, D ** -0.5)
    for n_bits, smooth in ((8, None), (6, (torch.rand(H * D, generator=g) + 0.5).half().cuda())):
        ref = ops.act_quant(o.view(1, -1, H * D), n_bits=n_bits, smooth=smooth)
        got = ops.attn_temporal_quant(qkv, B, T, S, H, D, D ** -0.5, n_bits=n_bits, smooth=smooth)
        torch.cuda.synchronize()
        assert torch.equal(got.codes, ref.codes) and torch.equal(got.delta, ref.delta)
        assert torch.equal(got.zp, ref.zp) and torch.equal(got.rowsum, ref.rowsum)
    assert ops.check_status() == 0


@pytest.mark.gpu
def test_static_quantiser_fused_forms_equal_the_two_pass_forms():
    """vq_gelu_act_quant_static / vq_ln_modulate_act_quant_static (the fused schedule of static checkpoints) against the
    same static quantiser applied to the separately formed tensor: GELU — torch's nn.GELU(approximate="tanh") in fp16, the op
    the reference runs — bit for bit; LayerNorm + modulate against the y of vq_ln_modulate_act_quant (two fp32 summation
    orders of the same LayerNorm: a last-bit difference of y may move a code by one step on a vanishing share of elements)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.nn.functional as F
    from viditq_b200 import ops
    g = torch.Generator().manual_seed(5)
    delta = torch.tensor([0.043], dtype=torch.float16, device="cuda")
    zp = torch.tensor([117.0], dtype=torch.float16, device="cuda")
    # ---- GELU, K = 4608 (the Mlp hidden tensor), values beyond the calibrated range saturate
    h = (torch.randn(1, 777, 4608, generator=g) * 2.5).half().cuda()
    a = ops.act_quant_static(h, delta, zp, gelu=True)
    b = ops.act_quant_static(F.gelu(h, approximate="tanh"), delta, zp)
    assert torch.equal(a.codes, b.codes) and torch.equal(a.rowsum, b.rowsum)
    assert int(a.codes.max()) == 255 and int(a.codes.min()) < 117
    # ---- LayerNorm + modulate, K = 1152, two modulation vectors (stacked cfg branches)
    x = (torch.randn(1, 2 * 300, 1152, generator=g) * 1.7 + 0.3).half().cuda()
    shift = (torch.randn(2, 1152, generator=g) * 0.2).half().cuda()
    scale = (torch.randn(2, 1152, generator=g) * 0.2).half().cuda()
    d2 = torch.tensor([0.021], dtype=torch.float16, device="cuda")
    z2 = torch.tensor([131.0], dtype=torch.float16, device="cuda")
    y = ops.ln_modulate_act_quant(x, shift, scale, want_y=True, rows_per_mod=300)[1]
    a = ops.act_quant_static(x, d2, z2, ln=(shift, scale), rows_per_mod=300)
    b = ops.act_quant_static(y, d2, z2)
    diff = (a.codes.int() - b.codes.int()).abs()
    frac = (diff > 0).float().mean().item()
    print("LN-fused static quantiser vs two passes: codes differing %.2e, largest step %d" % (frac, int(diff.max())))
    assert int(diff.max()) <= 1 and frac <= 1e-4
    assert ops.check_status() == 0
