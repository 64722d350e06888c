"""2-GPU test of the cfg-branch pair split (needs two CUDA devices; skipped otherwise): a pair of NCCL ranks running one
CFG branch each + the 2 MB exchange produces, on both ranks, the bit-identical latent of the same step run as two
forward calls on one GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _step_inputs(dev):
    g = torch.Generator().manual_seed(7)
    z = torch.randn(1, 4, 16, 32, 32, generator=g).to(dev)
    yc = torch.randn(1, 1, 120, 4096, generator=g).to(dev)
    yu = torch.randn(1, 1, 120, 4096, generator=g).to(dev)
    mask = torch.zeros(1, 120, dtype=torch.int64)
    mask[0, :77] = 1
    return z, yc, yu, mask.to(dev)


def _build(dev):
    import bench
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    torch.manual_seed(11)
    model = STDiT(input_size=(16, 32, 32), depth=2, hidden_size=1152, num_heads=16).eval()
    wq, aq = bench.quant_cfgs()
    aq["n_spatial_token"] = 16 * 16          # this test's latent is 32 x 32 -> 256 spatial tokens per frame
    qnn = QuantModel(model, wq, aq)
    qnn.cfg_split = True
    qnn.to(dev).half()
    model.dtype = torch.float16
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.fp_layer_list = bench.FP_LAYERS
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    return qnn, model


_STEP_CONST = {}


def _one_step(model, ops, ddim, z, y, mask, dev, exchange=None):
    if dev not in _STEP_CONST:       # host-side planning and H2D copies happen once, outside any graph capture
        plan = model.mask_select_plan(mask)
        _STEP_CONST[dev] = (torch.full((1,), float(ddim.model_timestep(ddim.num_timesteps - 1)), device=dev),
                            ddim.coefficients(ddim.num_timesteps - 1, "cpu").to(dev), plan, model.kv_segments(plan[1], dev))
    t, coef, plan, seg = _STEP_CONST[dev]
    if exchange is None:
        oc = model.forward_fused(z, t, y[0], plan=plan, segments=seg)
        ou = model.forward_fused(z, t, y[1], plan=plan, segments=seg)
    else:
        oc, ou = exchange(model.forward_fused(z, t, y, plan=plan, segments=seg))
    return ops.cfg_ddim_step(oc, ou, z, coef, ddim.cfg_scale)


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.set_grad_enabled(False)
    from viditq_b200 import ops, shard
    from viditq_b200.sampler import SpacedDDIM
    qnn, model = _build(dev)
    z, yc, yu, mask = _step_inputs(dev)
    grp = shard.cfg_pair_groups()
    ddim = SpacedDDIM(num_sampling_steps=100, cfg_scale=4.0)
    qnn.set_timestep_id_for_quantlayer(999.0)
    new = _one_step(model, ops, ddim, z, yu if shard.cfg_branch() else yc, mask, dev,
                    exchange=lambda o: shard.exchange_cfg_branches(o, grp))
    ref = _one_step(model, ops, ddim, z, (yc, yu), mask, dev) if rank == 0 else None
    # the same pair step as CUDA-graph segments around the eagerly issued all_gather (shard.SegmentedGraph)
    sg = shard.SegmentedGraph()
    seg_out = sg.capture(lambda: _one_step(model, ops, ddim, z, yu if shard.cfg_branch() else yc, mask, dev,
                                           exchange=lambda o: shard.exchange_cfg_branches(o, grp)))
    sg.replay()
    sg.replay()
    torch.cuda.synchronize()
    ret[rank] = (new.cpu().numpy(), None if ref is None else ref.cpu().numpy(), seg_out.cpu().numpy(), sg.counts())
    dist.destroy_process_group()


def test_cfg_branch_pair_equals_single_gpu_step():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    new0, ref, seg0, counts = ret[0]
    new1, _, seg1, _ = ret[1]
    import numpy as np
    assert np.isfinite(ref).all() and float(np.abs(ref).max()) > 0
    assert np.array_equal(new0, ref)       # the pair reproduces the single-GPU step bit for bit ...
    assert np.array_equal(new0, new1)      # ... on both ranks
    assert counts == (2, 1)                # graph | all_gather | graph
    assert np.array_equal(seg0, ref) and np.array_equal(seg1, ref)      # ... and so do the replayed graph segments


def _frames_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.set_grad_enabled(False)
    from viditq_b200 import shard
    qnn, model = _build(dev)
    z, yc, yu, mask = _step_inputs(dev)
    qnn.set_timestep_id_for_quantlayer(999.0)
    t = torch.full((2,), 999.0, device=dev)
    y2 = torch.cat([yc, yu])
    plan = model.mask_select_plan(mask.repeat(2, 1))
    seg = model.kv_segments(plan[1], dev)
    t0, t1 = shard.frame_slice(16)
    z_loc = z[:, :, t0:t1].contiguous()
    out = model.forward_fused(torch.cat([z_loc, z_loc]), t, y2, plan=plan, segments=seg, independent=True,
                              frames=(None, world, rank))
    ref = None
    if rank == 0:
        ref = model.forward_fused(torch.cat([z, z]), t, y2, plan=plan, segments=seg, independent=True)
    sg = shard.SegmentedGraph()
    seg_out = sg.capture(lambda: model.forward_fused(torch.cat([z_loc, z_loc]), t, y2, plan=plan, segments=seg,
                                                     independent=True, frames=(None, world, rank)))
    sg.replay()
    sg.replay()
    torch.cuda.synchronize()
    ret[rank] = (out.cpu().numpy(), (t0, t1), None if ref is None else ref.cpu().numpy(), seg_out.cpu().numpy(), sg.counts())
    dist.destroy_process_group()


def test_frame_sharded_forward_equals_single_gpu_forward():
    """16 frames over two NCCL ranks (8 each): every rank's slice of the model output is bit-identical to the same frames
    of the single-GPU forward — per-token quantisation, the frame-local kernels and the all-to-all of codes around the
    temporal attention change nothing in the arithmetic."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import numpy as np
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_frames_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    ref = ret[0][2]
    assert np.isfinite(ref).all() and float(np.abs(ref).max()) > 0
    for rank in range(2):
        out, (t0, t1), _, seg_out, counts = ret[rank]
        assert out.shape == ref[:, :, t0:t1].shape
        assert np.array_equal(out, ref[:, :, t0:t1]), rank
        assert counts == (2 * 2 + 1, 2 * 2)       # two exchanges per block: 5 graph segments around 4 all-to-alls
        assert np.array_equal(seg_out, ref[:, :, t0:t1]), rank   # replayed graph segments: same bits
