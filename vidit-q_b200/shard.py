"""Multi-GPU partitioning of the path: independent diffusion samples (prompts) sharded over ranks.

One process per GPU, a full weight replica each, no collective on the data path.  The only exchanges are the timing
barrier / max-over-ranks reduction and one all_gather of the final latents (1 MB per sample) before VAE decode.
Per-token activation statistics are pooled over the per-rank batch (quirk Q1), so an N-way sharded run equals the
reference executed once per prompt with batch_size = 1 — not one batch-N run.

A second, exact split for latency (strong scaling of ONE sample): with cfg_split the conditional and unconditional forwards
of a denoise step are independent model calls (iddpm/__init__.py:156-157), so a pair of ranks runs one branch each and
exchanges the model outputs (2 MB per rank per step) before both apply the identical CFG + DDIM update.  Nothing else
crosses ranks; the result is bit-identical to the two calls made on one GPU (tests/test_gpu_multi.py).

A third split, for one 16-frame video on P GPUs: frame (T) sharding.  Every per-token linear, LayerNorm, per-token
quantiser, the spatial attention and the cross attention are frame-local; only the temporal attention couples frames.  A
rank keeps T/P frames of the residual stream; per block the *quantised* input of attn_temp's q|k|v (u8 codes + 8 bytes of
scales per token — a quarter of the fp16 bytes, and per-token quantisation does not care where a token lives) goes through
one all-to-all into the (all frames, S/P positions) layout, the temporal q|k|v GEMM + attention + the projection's
quantiser run there, and a second all-to-all brings the codes back for the projection GEMM with the local residual.
Bit-identical to the single-GPU schedule (every kernel is row-, sequence- or sample-local).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def assign_samples(n_samples, world_size=None, rank=None):
    """Prompt i -> rank i mod P (round-robin keeps per-rank counts within one of each other)."""
    if world_size is None:
        world_size, rank = world()
    return list(range(rank, n_samples, world_size))


def gather_latents(local_latents, n_samples):
    """local_latents: [n_local, ...] for assign_samples(n_samples). Returns [n_samples, ...] in prompt order on every
    rank (ranks with fewer samples pad the exchange; padding is dropped)."""
    world_size, rank = world()
    if world_size == 1:
        return local_latents
    per_rank = (n_samples + world_size - 1) // world_size
    pad = per_rank - local_latents.shape[0]
    buf = local_latents
    if pad:
        buf = torch.cat([local_latents, local_latents.new_zeros((pad,) + tuple(local_latents.shape[1:]))], 0)
    out = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf.contiguous())
    full = local_latents.new_empty((n_samples,) + tuple(local_latents.shape[1:]))
    for r in range(world_size):
        idx = assign_samples(n_samples, world_size, r)
        if idx:
            full[torch.tensor(idx)] = out[r][:len(idx)]
    return full


def max_over_ranks(values, device):
    """Device-time reduction used by bench.py: every multi-GPU number is the max over ranks."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    world_size, _ = world()
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


# ------------------------------------------------------------------------------------------------ cfg-branch pairs
def cfg_branch(rank=None):
    """0 = conditional, 1 = unconditional branch of the pair this rank belongs to."""
    if rank is None:
        _, rank = world()
    return rank % 2


def cfg_pair_groups():
    """Process groups of consecutive rank pairs (2i, 2i+1); every rank must call this (collective). Returns this rank's
    group (None when running single-process)."""
    world_size, rank = world()
    if world_size == 1:
        return None
    if world_size % 2:
        raise ValueError("cfg-branch split needs an even number of ranks (two per sample)")
    mine = None
    for i in range(0, world_size, 2):
        g = dist.new_group([i, i + 1])
        if rank in (i, i + 1):
            mine = g
    return mine


def exchange_cfg_branches(out_local, group=None):
    """out_local: this rank's model output [n, c_out, ...] for its branch. Returns (out_cond, out_uncond), both complete
    on both ranks of the pair, in that order regardless of the caller's branch."""
    world_size, rank = world()
    if world_size == 1:
        raise ValueError("exchange_cfg_branches needs an initialised process group")
    both = torch.empty((2,) + tuple(out_local.shape), dtype=out_local.dtype, device=out_local.device)
    src = out_local.contiguous()
    parts = [both[0], both[1]]
    _collective(lambda: dist.all_gather(parts, src, group=group))
    return both[0], both[1]       # group rank 0 = even global rank = conditional branch


# ------------------------------------------------------------------------------------------------ frame (T) sharding
def frame_slice(T, world_size=None, rank=None):
    """Frames [t0, t1) owned by this rank (T must divide evenly: the reference's 16 frames over 2 / 4 / 8 / 16 ranks)."""
    if world_size is None:
        world_size, rank = world()
    if T % world_size:
        raise ValueError(f"{T} frames do not divide over {world_size} ranks")
    per = T // world_size
    return rank * per, (rank + 1) * per


class SegmentedGraph:
    """A step that contains torch.distributed collectives, captured as CUDA-graph SEGMENTS around eagerly issued NCCL
    calls: graph 0 | collective | graph 1 | collective | ...  The step function is run once under `capture`; every
    collective of this module, when it is reached, ends the running capture, is issued eagerly on the same stream (and
    remembered with its send / receive tensors, which are static: they come from the segments' shared memory pool) and a
    new capture begins behind it.  `replay` launches the segments and re-issues the collectives in order.  A frame-sharded
    forward (~700 kernel launches + 56 exchanges) becomes 57 graph launches + 56 NCCL calls; nothing about NCCL itself
    has to be capturable (capturing the collectives inside one whole-step graph deadlocked on the 2-GPU box)."""
    _active = None

    def __init__(self):
        self.items = []          # ("graph", CUDAGraph) | ("comm", callable)
        self.pool = None
        self.stream = None
        self._g = None

    def _begin(self):
        self._g = torch.cuda.CUDAGraph()
        self._g.capture_begin(pool=self.pool, capture_error_mode="thread_local")

    def _end(self):
        self._g.capture_end()
        self.items.append(("graph", self._g))
        self._g = None

    def capture(self, fn):
        """Run fn() once, recording it; returns fn's result (tensors of the last segment: static across replays)."""
        if SegmentedGraph._active is not None:
            raise RuntimeError("nested SegmentedGraph capture")
        self.pool = torch.cuda.graph_pool_handle()
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        SegmentedGraph._active = self
        try:
            with torch.cuda.stream(self.stream):
                self._begin()
                try:
                    out = fn()
                finally:
                    if self._g is not None:
                        self._end()
        finally:
            SegmentedGraph._active = None
        torch.cuda.current_stream().wait_stream(self.stream)
        return out

    def comm(self, fn):
        self._end()
        fn()
        self.items.append(("comm", fn))
        self._begin()

    def replay(self):
        for kind, it in self.items:
            if kind == "graph":
                it.replay()
            else:
                it()

    def counts(self):
        return (sum(1 for k, _ in self.items if k == "graph"), sum(1 for k, _ in self.items if k == "comm"))


def _collective(fn):
    """Issue a collective now; inside a SegmentedGraph capture it becomes a segment boundary."""
    if SegmentedGraph._active is not None:
        SegmentedGraph._active.comm(fn)
    else:
        fn()


def _all_to_all(send, group):
    recv = torch.empty_like(send)
    _collective(lambda: dist.all_to_all_single(recv, send, group=group))
    return recv


def frames_to_spatial(t, B, T_loc, S, P, group=None):
    """t: [B * T_loc * S, F] rows in (b, local frame, position) order on every rank -> [B * T * (S / P), F] rows in
    (b, frame, local position) order: all T = P * T_loc frames of this rank's S / P positions."""
    Sp, F = S // P, t.shape[-1]
    send = t.view(B, T_loc, P, Sp, F).permute(2, 0, 1, 3, 4).contiguous()          # [dst rank, b, t_loc, s', F]
    recv = _all_to_all(send, group)                                                # [src rank = frame chunk, b, t_loc, s', F]
    return recv.permute(1, 0, 2, 3, 4).reshape(B * P * T_loc * Sp, F)


def spatial_to_frames(t, B, T_loc, S, P, group=None):
    """Inverse of frames_to_spatial: [B * T * (S / P), F] -> [B * T_loc * S, F]."""
    Sp, F = S // P, t.shape[-1]
    send = t.view(B, P, T_loc, Sp, F).permute(1, 0, 2, 3, 4).contiguous()          # [dst rank = frame owner, b, t_loc, s', F]
    recv = _all_to_all(send, group)                                                # [src rank = position chunk, b, t_loc, s', F]
    return recv.permute(1, 2, 0, 3, 4).reshape(B * T_loc * S, F)


def _pack_meta(delta, zp, rowsum):
    """(delta fp16, zp fp16, rowsum i32) per row -> int32 [rows, 2] so that the scales travel in one exchange."""
    dz = torch.stack([delta, zp], dim=-1).contiguous().view(torch.int32)          # [rows, 1]
    return torch.cat([dz, rowsum.view(-1, 1)], dim=1).contiguous()


def _unpack_meta(meta):
    dz = meta[:, :1].contiguous().view(torch.float16)                              # [rows, 2]
    return dz[:, 0].contiguous(), dz[:, 1].contiguous(), meta[:, 1].contiguous()


def _exchange_act_codes_cuda(a, B, T_loc, S, P, to_spatial, group):
    """exchange_act_codes on the GPU: vq_row_pack writes the send buffer directly in rank-major order and unpacks the
    received rows into the target order — one pass each at HBM speed (the torch path below costs five strided byte copies
    per exchange: 8 of the 30 ms of a frame-sharded step at P = 2)."""
    from . import ops
    Sp = S // P
    n = B * T_loc * Sp                        # rows per (source, destination) rank pair
    if to_spatial:
        # local (b, t_loc, p, s') -> send [p, b, t_loc, s'];  recv [r, b, t_loc, s'] -> (b, frame = r T_loc + t_loc, s')
        send = ops.pack_rows(a, (B, T_loc, P, Sp), (T_loc * Sp, Sp, n, 1))
        recv = _all_to_all(send, group)
        return ops.unpack_rows(recv, a.K, (P, B, T_loc, Sp), (T_loc * Sp, P * T_loc * Sp, Sp, 1))
    # position-sharded (b, r, t_loc, s') -> send [r, b, t_loc, s'];  recv [p, b, t_loc, s'] -> local (b, t_loc, s = p Sp + s')
    send = ops.pack_rows(a, (B, P, T_loc, Sp), (T_loc * Sp, n, Sp, 1))
    recv = _all_to_all(send, group)
    return ops.unpack_rows(recv, a.K, (P, B, T_loc, Sp), (Sp, T_loc * S, S, 1))


def exchange_act_codes(a, B, T_loc, S, P, to_spatial, group=None):
    """Move per-token quantised activations (ops.ActCodes with one scale pair per row: G == 1) between the frame-sharded
    and the position-sharded layout.  ONE all-to-all: every row travels as K code bytes followed by its 8 bytes of
    (delta, zp, rowsum)."""
    if a.G != 1:
        raise ValueError("frame sharding moves per-token codes: batch-pooled statistics (G > 1) are not supported")
    if a.codes.is_cuda:
        return _exchange_act_codes_cuda(a, B, T_loc, S, P, to_spatial, group)
    fn = frames_to_spatial if to_spatial else spatial_to_frames
    K = a.codes.shape[-1]
    meta = _pack_meta(a.delta, a.zp, a.rowsum).view(torch.uint8).view(-1, 8)
    row = fn(torch.cat([a.codes.view(-1, K), meta], dim=1), B, T_loc, S, P, group)
    codes = row[:, :K].contiguous()
    delta, zp, rowsum = _unpack_meta(row[:, K:].contiguous().view(torch.int32).view(-1, 2))
    return type(a)(codes, delta, zp, rowsum, 1, codes.shape[0], a.K)
