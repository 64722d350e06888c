"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sample assignment, latent gather, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_samples, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from viditq_b200 import shard
    mine = shard.assign_samples(n_samples)
    local = torch.stack([torch.full((4, 2, 3), float(i)) for i in mine]) if mine else torch.zeros(0, 4, 2, 3)
    full = shard.gather_latents(local, n_samples)
    ok = all(bool((full[i] == float(i)).all()) for i in range(n_samples))
    t = shard.max_over_ranks([10.0 + rank, 5.0 - rank], "cpu")
    ret[rank] = (mine, ok, t)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [8, 5])
def test_sample_sharding_two_ranks(n_samples):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_samples, ret), nprocs=world, join=True)
    a, b = ret[0], ret[1]
    assert sorted(a[0] + b[0]) == list(range(n_samples)) and not set(a[0]) & set(b[0])
    assert abs(len(a[0]) - len(b[0])) <= 1
    assert a[1] and b[1]                       # every rank sees all latents in prompt order
    assert a[2] == b[2] == [11.0, 5.0]         # max over ranks


def test_single_process_is_identity():
    from viditq_b200 import shard
    assert shard.assign_samples(3) == [0, 1, 2]
    x = torch.randn(3, 2)
    assert shard.gather_latents(x, 3) is x
    assert shard.max_over_ranks([1.5], "cpu") == [1.5]


def _cfg_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from viditq_b200 import shard
    grp = shard.cfg_pair_groups()
    branch = shard.cfg_branch()
    pair = rank // 2
    out_local = torch.full((1, 8, 2, 3), 100.0 * pair + branch)     # what this rank's forward would return
    oc, ou = shard.exchange_cfg_branches(out_local, grp)
    ret[rank] = (branch, float(oc.flatten()[0]), float(ou.flatten()[0]))
    dist.destroy_process_group()


def test_cfg_branch_pairs_four_ranks():
    """Two samples x two CFG branches on four gloo ranks: each pair exchanges only inside the pair, cond first."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    world = 4
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cfg_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for rank in range(world):
        branch, oc, ou = ret[rank]
        assert branch == rank % 2
        assert (oc, ou) == (100.0 * (rank // 2), 100.0 * (rank // 2) + 1.0)


def _frames_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from viditq_b200 import shard
    B, T, S, F = 2, 4, 6, 3
    P = world
    t0, t1 = shard.frame_slice(T)
    T_loc = t1 - t0
    # global tensor g[b, t, s, f] = unique id; this rank holds frames [t0, t1)
    g = torch.arange(B * T * S * F, dtype=torch.int32).view(B, T, S, F)
    local = g[:, t0:t1].reshape(B * T_loc * S, F).contiguous()
    sp = shard.frames_to_spatial(local, B, T_loc, S, P)
    Sp = S // P
    want_sp = g[:, :, rank * Sp:(rank + 1) * Sp].reshape(B * T * Sp, F)
    back = shard.spatial_to_frames(sp, B, T_loc, S, P)

    class A:                                  # stand-in for ops.ActCodes (same constructor signature)
        def __init__(self, codes, delta, zp, rowsum, G, rows, K):
            self.codes, self.delta, self.zp, self.rowsum, self.G, self.rows, self.K = codes, delta, zp, rowsum, G, rows, K
    rows = B * T_loc * S
    ids = g[:, t0:t1, :, 0].reshape(rows)
    a = A(local.to(torch.uint8), (ids.float() * 0.5).half(), (ids % 7).half(), ids * 3, 1, rows, F)
    a_sp = shard.exchange_act_codes(a, B, T_loc, S, P, True)
    ids_sp = g[:, :, rank * Sp:(rank + 1) * Sp, 0].reshape(-1)
    meta_ok = (torch.equal(a_sp.delta, (ids_sp.float() * 0.5).half()) and torch.equal(a_sp.zp, (ids_sp % 7).half())
               and torch.equal(a_sp.rowsum, ids_sp * 3) and torch.equal(a_sp.codes, want_sp.to(torch.uint8)))
    a_fr = shard.exchange_act_codes(a_sp, B, T_loc, S, P, False)
    rt_ok = (torch.equal(a_fr.codes, a.codes) and torch.equal(a_fr.delta, a.delta) and torch.equal(a_fr.rowsum, a.rowsum))
    # the GPU path of the same exchange (vq_row_pack: rows packed straight into rank-major order, unpacked into the target
    # order) with the kernel restated in numpy (tests/cpu_ops.py): identical layouts and scales
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import cpu_ops
    from viditq_b200 import ops
    ops.pack_rows, ops.unpack_rows = cpu_ops.pack_rows, cpu_ops.unpack_rows
    K16 = 16
    wide = torch.arange(rows * K16, dtype=torch.int64).view(rows, K16).remainder(251).to(torch.uint8) + local[:, :1].to(torch.uint8)
    aw = ops.ActCodes(wide, a.delta, a.zp, a.rowsum.to(torch.int32), 1, rows, K16)
    g_sp = shard._exchange_act_codes_cuda(aw, B, T_loc, S, P, True, None)
    t_sp = shard.exchange_act_codes(A(wide, a.delta, a.zp, a.rowsum.to(torch.int32), 1, rows, K16), B, T_loc, S, P, True)
    g_fr = shard._exchange_act_codes_cuda(g_sp, B, T_loc, S, P, False, None)
    gpu_path_ok = (torch.equal(g_sp.codes, t_sp.codes) and torch.equal(g_sp.delta, t_sp.delta) and torch.equal(g_sp.zp, t_sp.zp)
                   and torch.equal(g_sp.rowsum, t_sp.rowsum) and torch.equal(g_fr.codes, wide) and torch.equal(g_fr.delta, a.delta)
                   and torch.equal(g_fr.rowsum, a.rowsum.to(torch.int32)))
    ret[rank] = (torch.equal(sp, want_sp), torch.equal(back, local), meta_ok, rt_ok, a_sp.rows == B * T * Sp, gpu_path_ok)
    dist.destroy_process_group()


def test_frame_sharding_layout_exchange_two_ranks():
    """frames <-> positions all-to-all (viditq_b200.shard) on gloo: every rank ends up with all frames of its S / P
    positions, the inverse restores the frame-sharded layout, and the per-token scales travel with their codes."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_frames_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for rank in range(world):
        assert all(ret[rank]), (rank, ret[rank])
