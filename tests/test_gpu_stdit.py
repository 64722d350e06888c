"""GPU model-level parity: viditq_b200 STDiT (2 blocks, hidden 1152, 4x16x16 latent, synthetic weights) with the
reference's own quant ckpt, against the W8A8 output of the unmodified reference QuantModel(STDiT) run in fp16
(tests/golden/stdit_small_golden.npz).  Tolerance of the north star: 1e-3 relative (stated below per norm)."""
import numpy as np
import pytest
import torch

from test_stdit_graph_cpu import FP_LAYERS, build_qnn, ckpt_from_golden, small  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qnn_gpu(small):  # noqa: F811
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    qnn, model = build_qnn(small)
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.set_quant_params_dict(ckpt_from_golden(small))
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.cuda()
    qnn.half()
    model.dtype = torch.float16
    return qnn, model


def _inputs(small):  # noqa: F811
    return (torch.from_numpy(small["x"]).cuda(), torch.from_numpy(small["t"]).cuda(),
            torch.from_numpy(small["y"]).cuda(), torch.from_numpy(small["mask"]).cuda())


def _set_w8a8(qnn):
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")


def _rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b).max() / np.abs(b).max(), np.linalg.norm(a - b) / np.linalg.norm(b)


def test_fp16_graph_on_gpu_matches_reference_fp16(qnn_gpu, small):  # noqa: F811
    qnn, model = qnn_gpu
    qnn.set_quant_state(False, False)
    x, t, y, mask = _inputs(small)
    with torch.no_grad():
        out = qnn(x, t, y, mask=mask).cpu().numpy()
    inf, l2 = _rel(out, small["out_fp16"])
    assert inf < 5e-3 and l2 < 2e-3, (inf, l2)      # fp16 graphs on different back ends (CPU eager vs GPU SDPA)


def test_w8a8_layerwise_matches_reference(qnn_gpu, small):  # noqa: F811
    from viditq_b200 import ops
    qnn, model = qnn_gpu
    _set_w8a8(qnn)
    x, t, y, mask = _inputs(small)
    n0 = ops.launch_count()
    with torch.no_grad():
        out = qnn(x, t, y, mask=mask).cpu().numpy()
    assert ops.launch_count() - n0 >= 2 * 13 * 2      # act-quant + GEMM per quantised linear, our kernels ran
    assert ops.check_status() == 0
    inf, l2 = _rel(out, small["out_w8a8"])
    print("layerwise vs reference W8A8: rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    # quantisation noise itself (W8A8 vs fp16) is 6.4e-3 rel-L2 on this model; parity must be far inside it
    assert l2 < 1e-3 and inf < 3e-3, (inf, l2)


def test_w8a8_fused_schedule_matches_layerwise_and_reference(qnn_gpu, small):  # noqa: F811
    qnn, model = qnn_gpu
    _set_w8a8(qnn)
    x, t, y, mask = _inputs(small)
    with torch.no_grad():
        ref_sched = qnn(x, t, y, mask=mask).cpu().numpy()
        qnn.set_timestep_id_for_quantlayer(float(small["t"][0]))
        fused = model.forward_fused(x, t, y, mask=mask).cpu().numpy()
    inf, l2 = _rel(fused, ref_sched)
    print("fused vs layerwise: rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert l2 < 3e-4, (inf, l2)                       # same rounding points; only LayerNorm statistics differ
    inf, l2 = _rel(fused, small["out_w8a8"])
    print("fused vs reference W8A8: rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert l2 < 1e-3 and inf < 3e-3, (inf, l2)
