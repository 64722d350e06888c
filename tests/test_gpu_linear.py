"""GPU parity of the ONE-LAUNCH fused QuantLinear (vq_linear_w8a8 -> vq_linear_fused_kernel) and of vq_col_absmax.

The fused kernel quantises the activation panel in its producer warps and never writes codes to HBM, so it is pinned
through its output: for every shape / epilogue / pooling / LayerNorm / smooth-quant / bit-width combination the result
must be BIT-IDENTICAL to the two-launch sequence (vq_act_quant | vq_ln_modulate_act_quant -> vq_gemm_w8a8), whose codes are
pinned bit-exact to the reference-generated vectors (tests/test_gpu_kernels.py) — same codes, same integer accumulator,
same epilogue arithmetic.  One case is also checked against the CPU oracle's integer form directly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(autouse=True)
def fused_policy():
    """The fused kernel is opt-in (quantise pass + GEMM is the measured-faster default): switch it on for these tests."""
    if not torch.cuda.is_available():
        yield
        return
    from viditq_b200 import ops
    ops.set_linear_fused_policy(1)
    yield
    ops.set_linear_fused_policy(0)


def _weight(N, K, n_bits=8, seed=0):
    from viditq_b200 import ops
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(N, K, generator=g) * 0.03).half().cuda()
    b = (torch.randn(N, generator=g) * 0.02).half().cuda()
    wf = w.float()
    mn, mx = wf.min(1)[0].clamp(max=0), wf.max(1)[0].clamp(min=0)
    d = ((mx - mn) / (2 ** n_bits - 1)).half()
    z = torch.round(-mn / d.float()).half()
    return w, b, d, z


CASES = [
    # (G, rows, N, epi, ln, smooth, n_bits)
    (1, 2048, 3456, 0, True, False, 8),      # PixArt-512 qkv: LayerNorm + modulate in the producer warps
    (1, 2048, 1152, 2, False, False, 8),     # proj: gated residual, in place
    (1, 2048, 4608, 1, True, False, 8),      # fc1 with the GELU epilogue
    (2, 1024, 1152, 0, False, False, 8),     # CFG pair pooled statistics (quirk Q1), interleaved panels
    (2, 1024, 3456, 0, True, False, 8),
    (4, 512, 1152, 2, False, False, 8),
    (1, 109, 2304, 0, False, False, 8),      # kv_linear: one ragged panel, single-CTA variant
    (1, 240, 2304, 0, False, True, 8),       # two prompts, smooth-quant scale
    (1, 300, 1152, 2, False, False, 8),      # ragged last panel + phantom pair partner (3 panels)
    (1, 128, 32, 0, False, False, 8),        # final_layer.linear of PixArt: N = 32 < one n-tile
    (1, 4096, 1152, 0, False, True, 6),      # 6-bit activations, smooth
    (1, 1024, 1152, 0, True, True, 4),       # 4-bit activations, LN + smooth
    (1, 8192, 1152, 2, False, False, 8),     # more panels than SM pairs / n_split = 1
]


@pytest.mark.parametrize("G,rows,N,epi,ln,smooth,n_bits", CASES)
def test_fused_linear_is_bit_identical_to_quant_plus_gemm(G, rows, N, epi, ln, smooth, n_bits):
    _need_gpu()
    from viditq_b200 import ops
    K = 1152
    M = G * rows
    assert ops.linear_launch_count(G, rows, K) == 1, "shape expected on the fused kernel"
    g = torch.Generator().manual_seed(rows + N + epi)
    x = torch.randn(G, rows, K, generator=g).half()
    x[..., 5] *= 12
    x[:, ::7, 100] -= 30
    x = x.cuda()
    w, b, d, z = _weight(N, K, seed=N)
    sm = (torch.rand(K, generator=g) + 0.5).half().cuda() if smooth else None
    pw = ops.prep_weight(w, d, z, n_bits=8, smooth=sm, bias=b)
    shift = scale = None
    rpm = rows
    if ln:
        n_mod = G
        if G == 1 and rows % 2 == 0:
            n_mod, rpm = 2, rows // 2                 # two stacked samples with their own modulation (cfg_split)
        shift = (torch.randn(n_mod, K, generator=g) * 0.1).half().cuda()
        scale = (torch.randn(n_mod, K, generator=g) * 0.1).half().cuda()
    res = gate = None
    rpg = 0
    if epi == 2:
        res = torch.randn(M, N, generator=g).half().cuda()
        gate = torch.randn(2 if M % 2 == 0 else 1, N, generator=g).half().cuda()
        rpg = M // gate.shape[0]
    # reference: two launches
    if ln:
        a, _ = ops.ln_modulate_act_quant(x, shift, scale, n_bits=n_bits, smooth=sm, rows_per_mod=rpm)
    else:
        a = ops.act_quant(x, n_bits=n_bits, smooth=sm)
    ref = ops.gemm_w8a8(a, pw, epi=epi, res=res, gate=gate, rows_per_gate=rpg)
    # fused: one launch (in place over a copy of the residual for the gated variant)
    out = res.clone() if epi == 2 else None
    n0 = ops.launch_count()
    got = ops.linear_w8a8(x, pw, n_bits=n_bits, smooth=sm, ln=(shift, scale) if ln else None, rows_per_mod=rpm if ln else None,
                          epi=epi, res=out, gate=gate, rows_per_gate=rpg, out=out)
    assert ops.launch_count() - n0 == 1
    torch.cuda.synchronize()
    assert ops.check_status() == 0
    bad = int((got.view(torch.int16) != ref.view(torch.int16)).sum())
    assert bad == 0, f"{bad} / {ref.numel()} elements differ from the two-launch path"


def test_fused_linear_against_the_cpu_oracle():
    _need_gpu()
    from oracle import qdiff_oracle as O
    from viditq_b200 import ops
    rng = np.random.default_rng(3)
    M, K, N = 384, 1152, 384
    x = rng.standard_normal((1, M, K)).astype(np.float16)
    x[..., 9] *= 15
    w = (rng.standard_normal((N, K)) * 0.03).astype(np.float16)
    b = (rng.standard_normal(N) * 0.02).astype(np.float16)
    wd, wz = O.weight_init_params(w.astype(np.float32), 8)
    wd, wz = wd.astype(np.float16), wz.astype(np.float16)
    dev = lambda v: torch.from_numpy(v).cuda()  # noqa: E731
    pw = ops.prep_weight(dev(w), dev(wd), dev(wz), bias=dev(b))
    y = ops.linear_w8a8(dev(x), pw).cpu().numpy().reshape(1, M, N)
    oa = O.dynamic_act_quant(x)
    wq = O.weight_quant(w, wd, wz)
    yo = O.quant_linear_int(oa["codes"], oa["delta"], oa["zp"], oa["rowsum"], wq["codes"], wd, wz, b)
    bad = float((y.view(np.uint16) != yo.view(np.uint16)).mean())
    assert bad <= 1e-5, bad          # last-bit double rounding of the oracle's float64 evaluation only


def test_large_or_unsupported_shapes_take_two_launches():
    _need_gpu()
    from viditq_b200 import ops
    ops.set_linear_fused_policy(-1, 8192)
    assert ops.linear_launch_count(1, 4096, 1152) == 1 and ops.linear_launch_count(1, 16384, 1152) == 2    # size threshold
    ops.set_linear_fused_policy(0)
    assert ops.linear_launch_count(1, 128, 1152) == 2            # the default policy: never
    ops.set_linear_fused_policy(1)
    assert ops.linear_launch_count(1, 2048, 4608) == 2           # K = 4608 panel does not fit shared memory
    assert ops.linear_launch_count(3, 128, 1152) == 2            # pooling group that does not divide a panel
    x = torch.randn(1, 2048, 4608).half().cuda()
    w, b, d, z = _weight(1152, 4608)
    pw = ops.prep_weight(w, d, z, bias=b)
    ref = ops.gemm_w8a8(ops.act_quant(x), pw)
    n0 = ops.launch_count()
    got = ops.linear_w8a8(x, pw)
    assert ops.launch_count() - n0 == 2
    assert torch.equal(got, ref)


@pytest.mark.parametrize("G,n,K,gelu", [(1, 1000, 1152, False), (16, 64, 1152, False), (2, 4096, 4608, True), (3, 77, 2304, False)])
def test_col_absmax_is_exact(G, n, K, gelu):
    _need_gpu()
    from viditq_b200 import ops
    g = torch.Generator().manual_seed(G * n)
    x = (torch.randn(G, n, K, generator=g) * 3).half().cuda()
    got = ops.col_absmax(x, gelu=gelu)
    src = torch.nn.functional.gelu(x, approximate="tanh") if gelu else x
    ref = src.abs().max(dim=-2)[0]
    if gelu:      # the fused GELU may differ from ATen's in the last fp16 bit on rare elements
        assert (got.float() - ref.float()).abs().max() <= 2e-3 * ref.float().abs().max()
    else:
        assert torch.equal(got, ref)


@pytest.mark.parametrize("G,rows,N,epi,ln", [(1, 2048, 3456, 0, True), (2, 1024, 1152, 2, False), (1, 109, 2304, 0, False),
                                              (1, 300, 1152, 1, False), (1, 4096, 4608, 0, True)])
def test_packed_int4_weights_equal_byte_codes(G, rows, N, epi, ln):
    """J2: W4A8 with the weight operand streamed as packed INT4 (vq_linear_w4a8: TMA of half-width tiles + converter warps)
    must be bit-identical to the same 4-bit codes stored one per byte (vq_linear_w8a8), and to quantise pass + GEMM."""
    _need_gpu()
    import os
    from viditq_b200 import ops
    K, M = 1152, G * rows
    g = torch.Generator().manual_seed(N + rows)
    x = torch.randn(G, rows, K, generator=g).half()
    x[..., 3] *= 9
    x = x.cuda()
    w = (torch.randn(N, K, generator=g) * 0.03).half().cuda()
    b = (torch.randn(N, generator=g) * 0.02).half().cuda()
    mn, mx = w.float().min(1)[0].clamp(max=0), w.float().max(1)[0].clamp(min=0)
    d = ((mx - mn) / 15).half()
    z = torch.round(-mn / d.float()).half()
    pw = ops.prep_weight(w, d, z, n_bits=4, bias=b)
    assert int(pw.codes.max()) <= 15
    shift = scale = None
    if ln:
        shift = (torch.randn(G, K, generator=g) * 0.1).half().cuda()
        scale = (torch.randn(G, K, generator=g) * 0.1).half().cuda()
    res = gate = None
    if epi == 2:
        res = torch.randn(M, N, generator=g).half().cuda()
        gate = torch.randn(1, N, generator=g).half().cuda()
    kw = dict(ln=(shift, scale) if ln else None, epi=epi, gate=gate, rows_per_gate=M if epi == 2 else 0)
    ref = ops.linear_w8a8(x, pw, res=res, **kw)                          # byte codes through the fused kernel
    if ln:
        a, _ = ops.ln_modulate_act_quant(x, shift, scale)
    else:
        a = ops.act_quant(x)
    two = ops.gemm_w8a8(a, pw, epi=epi, res=res, gate=gate, rows_per_gate=M if epi == 2 else 0)
    ops.pack_u4(pw)
    # the packed form really is two codes per byte
    pk = pw.packed.cpu().numpy()
    cd = pw.codes.cpu().numpy()
    assert pk.shape == (N, K // 2) and np.array_equal(pk & 15, cd[:, 0::2]) and np.array_equal(pk >> 4, cd[:, 1::2])
    n0 = ops.launch_count()
    got = ops.linear_w8a8(x, pw, res=res, **kw)                          # dispatches to vq_linear_w4a8
    assert ops.launch_count() - n0 == 1
    torch.cuda.synchronize()
    assert torch.equal(got, ref) and torch.equal(got, two)


OVERLAP_CASES = [
    # (rows, N, epi, ln, smooth, n_bits)
    (2048, 3456, 0, True, False, 8),
    (4096, 1152, 2, False, False, 8),        # gated residual, in place
    (16384, 4608, 0, True, True, 8),         # fc1 shape, LN + smooth
    (1300, 1152, 2, False, False, 8),        # ragged last m-panel
    (32768, 1152, 0, False, False, 6),       # stacked cfg_split size, 6-bit activations
    (32768, 3456, 0, True, False, 8),
]


@pytest.mark.parametrize("rows,N,epi,ln,smooth,n_bits", OVERLAP_CASES)
def test_overlapped_quantiser_gemm_is_bit_identical(rows, N, epi, ln, smooth, n_bits):
    """Policy mode 2: the quantise pass runs INSIDE the persistent GEMM (two quantiser warpgroups ahead of the MMAs, codes
    through an L2-resident scratch, per-panel ready flags).  Same quantiser code, same GEMM: output bit-identical to the
    two-launch sequence."""
    _need_gpu()
    from viditq_b200 import ops
    K, M = 1152, rows
    g = torch.Generator().manual_seed(rows + N)
    x = torch.randn(1, rows, K, generator=g).half()
    x[..., 5] *= 12
    x = x.cuda()
    w, b, d, z = _weight(N, K, seed=N)
    sm = (torch.rand(K, generator=g) + 0.5).half().cuda() if smooth else None
    pw = ops.prep_weight(w, d, z, n_bits=8, smooth=sm, bias=b)
    shift = scale = None
    rpm = rows
    if ln:
        if rows % 2 == 0:
            rpm = rows // 2
        shift = (torch.randn(rows // rpm, K, generator=g) * 0.1).half().cuda()
        scale = (torch.randn(rows // rpm, K, generator=g) * 0.1).half().cuda()
    res = gate = None
    rpg = 0
    if epi == 2:
        res = torch.randn(M, N, generator=g).half().cuda()
        gate = torch.randn(1, N, generator=g).half().cuda()
        rpg = M
    if ln:
        a, _ = ops.ln_modulate_act_quant(x, shift, scale, n_bits=n_bits, smooth=sm, rows_per_mod=rpm)
    else:
        a = ops.act_quant(x, n_bits=n_bits, smooth=sm)
    ref = ops.gemm_w8a8(a, pw, epi=epi, res=res, gate=gate, rows_per_gate=rpg)
    out = res.clone() if epi == 2 else None
    ops.set_linear_fused_policy(2)
    try:
        assert ops.linear_launch_count(1, rows, K) == 1
        for _ in range(3):       # repeated launches reuse the scratch / flags
            if epi == 2:
                out.copy_(res)
            got = ops.linear_w8a8(x, pw, n_bits=n_bits, smooth=sm, ln=(shift, scale) if ln else None,
                                  rows_per_mod=rpm if ln else None, epi=epi, res=out, gate=gate, rows_per_gate=rpg, out=out)
        torch.cuda.synchronize()
    finally:
        ops.set_linear_fused_policy(1)
    assert ops.check_status() == 0
    bad = int((got.view(torch.int16) != ref.view(torch.int16)).sum())
    assert bad == 0, f"{bad} / {ref.numel()} elements differ from the two-launch path"
