"""The glue oracle (oracle/glue_oracle.py: softmax attention, patch embedding) against outputs of the unmodified reference
classes (tests/golden/glue_golden.npz), and — on the GPU — the kernels against the oracle on the same vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import glue_oracle as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(ROOT, "tests", "golden", "glue_golden.npz"))
    return {k: z[k] for k in z.files}


def test_attention_oracle_matches_reference_eager_attention(gold):
    out = G.attention(gold["q"], gold["k"], gold["v"], int(gold["heads"]))
    err = np.abs(out - gold["core"]).max() / np.abs(gold["core"]).max()
    assert err < 2e-6, err


def test_cross_attention_oracle_is_blockwise_attention(gold):
    q = np.concatenate([gold["q"], gold["k"]], 0)                  # two "samples"
    kv = np.concatenate([gold["k"][0, :50], gold["v"][0, :50]], -1)
    kv = np.concatenate([kv, np.concatenate([gold["v"][0, :7], gold["q"][0, :7]], -1)], 0)   # prompts of 50 and 7 rows
    out = G.cross_attention(q, kv, [50, 7], int(gold["heads"]))
    ref0 = G.attention(q[:1], gold["k"][:, :50], gold["v"][:, :50], int(gold["heads"]))
    ref1 = G.attention(q[1:], gold["v"][:, :7], gold["q"][:, :7], int(gold["heads"]))
    assert np.array_equal(out[0], ref0[0]) and np.array_equal(out[1], ref1[0])


def test_patch_embed_oracle_matches_reference_patchembed3d(gold):
    out = G.patch_embed(gold["z"], gold["pe_w"], gold["pe_b"], None, (2, 2), fp16_graph=False)
    err = np.abs(out - gold["pe_out"]).max() / np.abs(gold["pe_out"]).max()
    assert out.shape == gold["pe_out"].shape and err < 2e-6, err


@pytest.mark.gpu
def test_attention_kernels_match_the_oracle_on_reference_vectors(gold):
    """Temporal kernel (16 keys) and tcgen05 spatial kernel (256-token sequences built from the golden q/k/v rows) against
    the oracle: fp16 inputs, fp16 output rounding and fp16 P are the only differences."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200 import ops
    H, D = int(gold["heads"]), 72
    C = H * D
    rng = np.random.default_rng(0)
    rows = rng.integers(0, gold["q"].shape[1], size=512)
    q, k, v = (gold[n][0, rows].astype(np.float16) for n in ("q", "k", "v"))       # [512, C] = 2 sequences of 256
    qkv = torch.from_numpy(np.stack([q, k, v], 1).reshape(512, 3 * C)).cuda()
    out = ops.attn_spatial(qkv, 2, 256, H, D, D ** -0.5).cpu().numpy().astype(np.float32)
    ref = G.attention(q.reshape(2, 256, C), k.reshape(2, 256, C), v.reshape(2, 256, C), H).reshape(512, C)
    assert np.abs(out - ref).max() <= 4e-3 * np.abs(ref).max() + 1e-3
    # temporal: B=1, T=16, S=32 in the (T S) layout
    out_t = ops.attn_temporal(qkv, 1, 16, 32, H, D, D ** -0.5).cpu().numpy().astype(np.float32)
    q5, k5, v5 = (a.reshape(16, 32, C).transpose(1, 0, 2) for a in (q, k, v))      # [S, T, C]: one sequence per position
    ref_t = G.attention(q5, k5, v5, H).transpose(1, 0, 2).reshape(512, C)
    assert np.abs(out_t - ref_t).max() <= 4e-3 * np.abs(ref_t).max() + 1e-3


@pytest.mark.gpu
def test_patch_embed_kernel_matches_the_oracle_on_reference_vectors(gold):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viditq_b200 import ops
    z, w, b = gold["z"], gold["pe_w"], gold["pe_b"]
    S = (z.shape[3] // 2) * (z.shape[4] // 2)
    pos = (np.random.default_rng(1).standard_normal((S, w.shape[0])) * 0.5).astype(np.float16)
    out = ops.patch_embed(torch.from_numpy(z).cuda(), torch.from_numpy(w).half().cuda(), torch.from_numpy(b).half().cuda(),
                          torch.from_numpy(pos).cuda(), (2, 2)).cpu().numpy()
    ref = G.patch_embed(z, w, b, pos, (2, 2))
    assert out.shape == ref.shape
    # bit-identical except where the summation order of the 16-term dot product moves a rounding
    assert float((out.view(np.uint16) != ref.view(np.uint16)).mean()) < 2e-2
    assert np.abs(out.astype(np.float32) - ref.astype(np.float32)).max() <= 2e-3 * np.abs(ref.astype(np.float32)).max()
