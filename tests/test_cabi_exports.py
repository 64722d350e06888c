"""CPU-side check of the drop-in boundary: libviditq_b200.so builds for sm_100a, loads, and exports every entry point
include/viditq_b200.h declares (no compute call is made here — there is no GPU in this container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "viditq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(vq_[a-z0-9_]+)\s*\(", hdr)))


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
