"""CPU-side check of the drop-in boundary: libviditq_b200.so builds for sm_100a, loads, and exports every entry point
include/viditq_b200.h declares (no compute call is made here — there is no GPU in this container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "viditq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t)\s+(vq_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ["vq_prep_weight", "vq_act_quant", "vq_ln_modulate_act_quant", "vq_gemm_w8a8"]:
        assert must in names


def test_library_builds_loads_and_exports_every_declared_symbol():
    import viditq_b200
    path = viditq_b200.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.vq_version.restype = ctypes.c_int
    assert lib.vq_version() >= 100


def test_python_binding_lists_the_same_symbols():
    from viditq_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The GEMM must be a tcgen05/TMA kernel, not a recompiled mma.sync one (B200_PROFILING.md SASS table)."""
    import shutil
    import subprocess
    import viditq_b200
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", viditq_b200.build_library()], capture_output=True, text=True).stdout
    assert "UTCIMMA" in sass       # tcgen05.mma.kind::i8
    assert "UTMALDG" in sass       # cp.async.bulk.tensor loads
    assert "UTMASTG" in sass       # TMA stores from the epilogue
    assert "UTCIMMA.2CTA" in sass  # cta_group::2 CTA-pair variant
    assert "UTCBAR.2CTA.MULTICAST" in sass   # tcgen05.commit multicast to both CTAs of a pair
    assert "LDTM" in sass          # tcgen05.ld
    assert "UTCHMMA" in sass       # tcgen05.mma.kind::f16 (vq_attn_spatial: S = Q K^T and O += P V)
    assert "UTMALDG.4D" in sass    # 4-D q|k|v tensor maps of the attention
    assert "STTM" in sass          # tcgen05.st: P written to TMEM as the A operand of P V
    assert "IMMA." not in sass.replace("UTCIMMA", "")   # no legacy mma.sync integer path


def test_argument_validation_returns_error_codes_without_touching_the_device():
    """Error behaviour of the boundary: bad pointers / shapes are refused with VQ_ERR_ARG (-1) or VQ_ERR_UNSUPPORTED (-5)
    before any CUDA call is made, so this runs on a machine without a GPU."""
    import viditq_b200
    lib = ctypes.CDLL(viditq_b200.build_library())
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.vq_attn_spatial.argtypes = [vp, vp, i32, i32, i32, i32, f32, vp]
    lib.vq_attn_cross.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, f32, vp]
    lib.vq_patch_embed.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.vq_add_act_quant.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.vq_gemm_w8a8.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp]
    fake = 0x10000   # never dereferenced: every call below is rejected by its argument checks
    assert lib.vq_attn_spatial(None, fake, 1, 1024, 16, 72, 0.1, None) == -1          # null q|k|v
    assert lib.vq_attn_spatial(fake, fake, 1, 1000, 16, 72, 0.1, None) == -5          # S not a multiple of 256
    assert lib.vq_attn_spatial(fake, fake, 1, 1024, 16, 64, 0.1, None) == -5          # head_dim != 72
    assert lib.vq_attn_spatial(fake + 2, fake, 1, 1024, 16, 72, 0.1, None) == -1      # misaligned tensor
    assert lib.vq_attn_cross(fake, fake, fake, None, fake, 1, 256, 16, 72, 120, 120, 0.1, None) == -1   # null kv_start
    assert lib.vq_attn_cross(fake, fake, fake, fake, fake, 1, 256, 16, 72, 200, 200, 0.1, None) == -5   # prompt > 128 rows
    assert lib.vq_patch_embed(None, fake, None, None, 1, 4, 16, 64, 64, 2, 2, 1152, fake, None) == -1
    assert lib.vq_patch_embed(fake, fake, None, None, 1, 4, 16, 64, 64, 4, 4, 1152, fake, None) == -5   # 64-element patches
    assert lib.vq_patch_embed(fake, fake, None, None, 1, 4, 16, 63, 64, 2, 2, 1152, fake, None) == -5   # ragged grid
    assert lib.vq_add_act_quant(fake, None, 1024, 16, 1, 64, 1152, None, 8, fake, fake, fake, fake, None, None) == -1
    assert lib.vq_add_act_quant(fake, fake, 1024, 16, 1, 64, 4608, None, 8, fake, fake, fake, fake, None, None) == -5
    assert lib.vq_gemm_w8a8(fake, fake, fake, fake, 64, fake, fake, 64, 100, 1152, 0, None, 0, None, 0, fake, 100, None) == -1  # N % 8
    assert lib.vq_gemm_w8a8(fake, fake, fake, fake, 64, fake, fake, 64, 1152, 1152, 2, None, 0, None, 0, fake, 1152, None) == -1  # residual epilogue without res


def test_int8_attention_boundary_without_a_device():
    """vq_attn_i8_*: the workspace size the library reports equals the layout the host side mirrors (tests read codes and
    scales through it), unsupported shapes answer -1 / VQ_ERR_UNSUPPORTED, bad pointers VQ_ERR_ARG — no CUDA call is made."""
    import viditq_b200
    from viditq_b200 import ops
    lib = ctypes.CDLL(viditq_b200.build_library())
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.vq_attn_i8_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.vq_attn_i8_workspace_bytes.restype = i64
    lib.vq_attn_spatial_i8.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, vp]
    lib.vq_attn_i8_quantise.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.vq_attn_i8_attend.argtypes = [vp, vp, i32, i32, i32, i32, f32, vp]
    for n_seq, S, H in [(1, 256, 1), (3, 1024, 16), (32, 1024, 16), (2, 4096, 16)]:
        assert lib.vq_attn_i8_workspace_bytes(n_seq, S, H, 72) == ops.attn_i8_workspace_layout(n_seq, S, H)["total"]
    assert lib.vq_attn_i8_workspace_bytes(1, 384, 16, 72) == -1      # S not a multiple of 256
    assert lib.vq_attn_i8_workspace_bytes(1, 8192, 16, 72) == -1     # more than 64 key blocks
    assert lib.vq_attn_i8_workspace_bytes(1, 1024, 16, 64) == -1     # head_dim != 72
    fake = 0x10000
    assert lib.vq_attn_spatial_i8(None, fake, fake, 1, 1024, 16, 72, 0.1, None) == -1
    assert lib.vq_attn_spatial_i8(fake, fake, fake + 16, 1, 1024, 16, 72, 0.1, None) == -1    # workspace not 256-byte aligned
    assert lib.vq_attn_spatial_i8(fake, fake, fake, 1, 1000, 16, 72, 0.1, None) == -5
    assert lib.vq_attn_i8_quantise(fake, None, 1, 1024, 16, 72, None) == -1
    assert lib.vq_attn_i8_attend(fake, fake, 1, 1024, 16, 80, 0.1, None) == -5
