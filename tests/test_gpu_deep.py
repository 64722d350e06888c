"""Full-depth GPU parity: the numbers the headline rests on.

  * 28-block STDiT-XL/2 graph on a small latent against the W8A8 output of the UNMODIFIED reference at the same depth
    (tests/golden/stdit_deep_golden.npz, generator make_golden_deep.py), for both schedules (`forward` = one QuantLayer
    call per linear, `forward_fused` = the B200 schedule);
  * 5 DDIM steps through viditq_b200.sampler against the latents the reference scheduler produced after every step
    (IDDPM.ddim_sample_loop around forward_with_cfg, cfg_split);
  * at the benchmark's own size (16x512x512, 28 blocks): the stacked fused step against the layer-by-layer schedule.

End-to-end distances of a re-quantising network are stated next to the reference's OWN noise band at that depth — its fp16
simulation against the same simulation with identical codes but un-rounded dequantised operands, measured here on the
same GPU — and next to the quantisation error itself; per-layer parity (<= 1e-3, the north-star
tolerance) is asserted teacher-forced on all 364 quantised linears.
"""
import os

import numpy as np
import pytest
import torch

from test_stdit_graph_cpu import FP_LAYERS, quant_cfgs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / np.abs(b).max(), np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def deep():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    z = np.load(os.path.join(ROOT, "tests", "golden", "stdit_deep_golden.npz"))
    g = {k: z[k] for k in z.files}
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    torch.set_grad_enabled(False)
    model = STDiT(input_size=(4, 16, 16), depth=28)
    model.init_synthetic(seed=0)
    model.eval()
    wq, aq = quant_cfgs(int(g["T"]), int(g["S"]))
    qnn = QuantModel(model, wq, aq)
    qnn.cfg_split = True
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.init_weight_quant_params()      # fp32 on the CPU: make_golden_deep.py asserts this equals the reference's ckpt
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.cuda()
    qnn.half()
    model.dtype = torch.float16
    return g, qnn, model


def _set_w8a8(qnn):
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")


def _sim_forward(qnn, x, t, y, mask, exact):
    """The reference's simulated path on this GPU (oracle.torch_fake_quant, pinned bit-exact to the reference on CPU)."""
    from oracle import torch_fake_quant as TF
    saved = {}

    def make(layer):
        def fwd(inp, *a, **k):
            if not (layer.weight_quant and layer.act_quant):
                return saved[layer](inp)
            G, rows = layer._pool_view(inp)
            wq = layer.weight_quantizer
            out = TF.quant_linear_fake(inp.reshape(G, rows, inp.shape[-1]), layer.weight, layer.bias, wq.delta,
                                       wq.zero_point, wq.n_bits, layer.act_quantizer.n_bits, exact=exact)
            return out.reshape(*inp.shape[:-1], -1)
        return fwd
    for _, layer in qnn.quant_layers():
        saved[layer] = layer.forward
        layer.forward = make(layer)
    try:
        return qnn(x, t, y, mask=mask).cpu().numpy()
    finally:
        for layer in saved:
            del layer.forward


def test_28_block_forward_against_the_reference(deep):
    g, qnn, model = deep
    from viditq_b200 import ops
    x, t = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()
    y, mask = torch.from_numpy(g["y"]).cuda(), torch.from_numpy(g["mask"]).cuda()
    qnn.set_quant_state(False, False)
    floor = _rel(qnn(x, t, y, mask=mask).cpu().numpy(), g["out_fp16"])
    _set_w8a8(qnn)
    sim = _sim_forward(qnn, x, t, y, mask, exact=False)
    sim_exact = _sim_forward(qnn, x, t, y, mask, exact=True)
    band = _rel(sim, sim_exact)
    sim_vs_ref = _rel(sim, g["out_w8a8"])
    n0 = ops.launch_count()
    out = qnn(x, t, y, mask=mask).cpu().numpy()
    launches = ops.launch_count() - n0
    qnn.set_timestep_id_for_quantlayer(float(g["t"][0]))
    fused = model.forward_fused(x, t, y, mask=mask).cpu().numpy()
    assert ops.check_status() == 0 and np.isfinite(out).all() and np.isfinite(fused).all()
    a, b = _rel(out, g["out_w8a8"]), _rel(fused, g["out_w8a8"])
    c, d = _rel(out, sim_exact), _rel(fused, sim_exact)
    s = _rel(fused, out)
    qerr = float(g["quant_err_forward"])
    print("\n28 blocks, 256 tokens, vs the reference executed on CPU (fp16):")
    print("  cross-back-end fp16 floor (no quantisation)        rel-inf %.3e  rel-L2 %.3e" % floor)
    print("  reference simulation re-run on this GPU vs golden  rel-inf %.3e  rel-L2 %.3e" % sim_vs_ref)
    print("  reference noise band (fp16 sim vs exact operands)  rel-inf %.3e  rel-L2 %.3e" % band)
    print("  forward  (hook schedule, %4d launches) vs golden   rel-inf %.3e  rel-L2 %.3e" % ((launches,) + a))
    print("  forward_fused                        vs golden     rel-inf %.3e  rel-L2 %.3e" % b)
    print("  forward / forward_fused vs exact-operand sim       rel-L2 %.3e / %.3e" % (c[1], d[1]))
    print("  forward_fused vs forward                           rel-inf %.3e  rel-L2 %.3e" % s)
    print("  quantisation error itself (W8A8 vs fp16)           rel-L2 %.3e" % qerr)
    # At 28 blocks the reference's own simulation does not reproduce itself across back ends any better than this: moved
    # from the CPU to this GPU it lands 7.5e-3 from its golden output, and 7.4e-3 from itself without the fp16 rounding of
    # the dequantised operands (B200 run, r02) — three quarters of the 8-bit quantisation error.  The integer kernels must
    # sit inside that band (x1.25) around the reference, around the simulation and around each other, and inside the
    # quantisation error; the 1e-3 bar is asserted where it is a property of the arithmetic: per layer, below.
    lim = 1.25 * max(band[1], sim_vs_ref[1])
    for r in (a, b, c, d, s):
        assert r[1] <= lim and r[1] < qerr, (r, lim, qerr)


def test_per_layer_parity_at_full_depth(deep):
    """The north-star tolerance on every one of the 364 quantised linears, teacher-forced inside the 28-block model."""
    g, qnn, model = deep
    from oracle import torch_fake_quant as TF
    _set_w8a8(qnn)
    x, t = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()
    y, mask = torch.from_numpy(g["y"]).cuda(), torch.from_numpy(g["mask"]).cuda()
    saved, errs = {}, {}

    def make(name, layer):
        def fwd(inp, *a, **k):
            out = saved[layer](inp)
            if layer.weight_quant and layer.act_quant:
                G, rows = layer._pool_view(inp)
                wq = layer.weight_quantizer
                ref = TF.quant_linear_fake(inp.reshape(G, rows, inp.shape[-1]), layer.weight, layer.bias, wq.delta,
                                           wq.zero_point, wq.n_bits, layer.act_quantizer.n_bits).reshape(out.shape)
                dd, rr = out.float() - ref.float(), ref.float()
                errs[name] = ((dd.abs().max() / rr.abs().max()).item(), (dd.norm() / rr.norm()).item())
            return out
        return fwd
    for name, layer in qnn.quant_layers():
        saved[layer] = layer.forward
        layer.forward = make(name, layer)
    try:
        qnn(x, t, y, mask=mask)
    finally:
        for layer in saved:
            del layer.forward
    assert len(errs) == 13 * 28
    wi = max(errs.items(), key=lambda kv: kv[1][0])
    wl = max(errs.items(), key=lambda kv: kv[1][1])
    print("\nper-layer parity over %d layers: worst rel-inf %.3e (%s), worst rel-L2 %.3e (%s)"
          % (len(errs), wi[1][0], wi[0], wl[1][1], wl[0]))
    assert wi[1][0] <= 1e-3 and wl[1][1] <= 1e-3, (wi, wl)


def test_five_ddim_steps_against_the_reference_scheduler(deep):
    g, qnn, model = deep
    from functools import partial
    from viditq_b200.sampler import SpacedDDIM, ddim_sample_loop
    n_steps = int(g["n_steps"])
    ddim = SpacedDDIM(num_sampling_steps=n_steps, cfg_scale=float(g["cfg_scale"]))
    assert list(ddim.timestep_map) == list(g["timestep_map"])
    z0 = torch.from_numpy(g["z0"]).cuda()
    y, y_null = torch.from_numpy(g["y"]).cuda(), torch.from_numpy(g["y_null"]).cuda()
    mask = torch.from_numpy(g["mask"]).cuda()

    def run(forward, stacked=None):
        traj = []
        ddim_sample_loop(ddim, forward, z0.clone(), y, y_null, mask, qnn=qnn, on_step=lambda i, z: traj.append(z.cpu().numpy()),
                         stacked_forward=stacked)
        return np.stack(traj)
    qnn.set_quant_state(False, False)
    fp = run(lambda x, t, yy, mask=None: qnn(x, t, yy, mask=mask))
    _set_w8a8(qnn)
    # the reference's own simulated arithmetic (oracle.torch_fake_quant) run through the same 5 steps on THIS GPU: how far
    # the latents of the reference move when only the back end changes
    from oracle import torch_fake_quant as TF
    saved = {}

    def sim_layer(layer):
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
        layer.forward = sim_layer(layer)
    try:
        sim = run(lambda x, t, yy, mask=None: qnn(x, t, yy, mask=mask))
    finally:
        for layer in saved:
            del layer.forward
    hook = run(lambda x, t, yy, mask=None: qnn(x, t, yy, mask=mask))
    fused = run(lambda x, t, yy, mask=None: model.forward_fused(x, t, yy, mask=mask),
                stacked=partial(model.forward_fused, independent=True))
    print("\n5 DDIM steps (cfg_split, cfg_scale 4) vs the latents of the reference scheduler, rel-L2 per step:")
    print("  fp16 (no quantisation, back-end floor)  " + "  ".join("%.2e" % _rel(fp[k], g["traj_fp16"][k])[1] for k in range(n_steps)))
    print("  reference simulation on this GPU        " + "  ".join("%.2e" % _rel(sim[k], g["traj_w8a8"][k])[1] for k in range(n_steps)))
    print("  hook schedule (forward)                 " + "  ".join("%.2e" % _rel(hook[k], g["traj_w8a8"][k])[1] for k in range(n_steps)))
    print("  fused schedule (stacked forward_fused)  " + "  ".join("%.2e" % _rel(fused[k], g["traj_w8a8"][k])[1] for k in range(n_steps)))
    print("  quantisation error of the reference     " + "  ".join("%.2e" % _rel(g["traj_w8a8"][k], g["traj_fp16"][k])[1] for k in range(n_steps)))
    print("  fused vs hook schedule                  " + "  ".join("%.2e" % _rel(fused[k], hook[k])[1] for k in range(n_steps)))
    # CFG (u + 4 (c - u)) amplifies whatever separates two forwards, quantisation error and re-quantisation noise alike: the
    # kernels must stay inside the distance the reference's own simulation shows when it only changes back end (x1.25),
    # and inside the quantisation error of the reference's sampling run
    for k in range(n_steps):
        q_k = _rel(g["traj_w8a8"][k], g["traj_fp16"][k])[1]
        lim = 1.25 * _rel(sim[k], g["traj_w8a8"][k])[1]
        for tr in (hook, fused):
            assert np.isfinite(tr[k]).all()
            r = _rel(tr[k], g["traj_w8a8"][k])[1]
            assert r <= lim and r < q_k, (k, r, lim, q_k)


def test_benchmark_size_fused_step_equals_layerwise_schedule():
    """16x512x512, 28 blocks — the configuration of the headline number: forward_fused on the stacked cond | uncond pair
    (what bench.py times) against forward(), the reference's schedule of one QuantLayer call per linear, called twice."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import bench
    from viditq_b200 import ops
    torch.set_grad_enabled(False)
    torch.manual_seed(1234)
    dev = torch.device("cuda", 0)
    qnn, model = bench.build_model(dev, bench.DEPTH)
    g = torch.Generator().manual_seed(99)
    z = torch.randn(1, 4, bench.T_FRAMES, 64, 64, generator=g).to(dev)
    yc = torch.randn(1, 1, bench.PROMPT_LEN, 4096, generator=g).to(dev)
    yu = torch.randn(1, 1, bench.PROMPT_LEN, 4096, generator=g).to(dev)
    mask = torch.zeros(1, bench.PROMPT_LEN, dtype=torch.int64)
    mask[0, :109] = 1
    mask = mask.to(dev)
    t = torch.full((1,), 999.0, device=dev)
    qnn.set_timestep_id_for_quantlayer(999.0)
    both = model.forward_fused(torch.cat([z, z]), t.expand(2), torch.cat([yc, yu]), mask=mask, independent=True)
    torch.cuda.synchronize()
    ref_c = qnn(z, t, yc, mask=mask)
    ref_u = qnn(z, t, yu, mask=mask)
    qnn.set_quant_state(False, False)
    fp_c = qnn(z, t, yc, mask=mask)
    assert ops.check_status() == 0
    ref = torch.cat([ref_c, ref_u]).cpu().numpy()
    got = both.cpu().numpy()
    assert np.isfinite(got).all()
    inf, l2 = _rel(got, ref)
    qerr = _rel(ref_c.cpu().numpy(), fp_c.cpu().numpy())[1]
    print("\n16x512x512, 28 blocks: stacked forward_fused vs forward (hook schedule) rel-inf %.3e rel-L2 %.3e; "
          "quantisation error %.3e" % (inf, l2, qerr))
    assert l2 < qerr and l2 <= 1e-2, (inf, l2, qerr)
