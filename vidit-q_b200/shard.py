"""Multi-GPU partitioning of the path: independent diffusion samples (prompts) sharded over ranks.

One process per GPU, a full weight replica each, no collective on the data path.  The only exchanges are the timing
barrier / max-over-ranks reduction and one all_gather of the final latents (1 MB per sample) before VAE decode.
Per-token activation statistics are pooled over the per-rank batch (quirk Q1), so an N-way sharded run equals the
reference executed once per prompt with batch_size = 1 — not one batch-N run.

A second, exact split for latency (strong scaling of ONE sample): with cfg_split the conditional and unconditional forwards
of a denoise step are independent model calls (iddpm/__init__.py:156-157), so a pair of ranks runs one branch each and
exchanges the model outputs (2 MB per rank per step) before both apply the identical CFG + DDIM update.  Nothing else
crosses ranks; the result is bit-identical to the two calls made on one GPU (tests/test_gpu_multi.py).
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
    parts = [torch.empty_like(out_local) for _ in range(2)]
    dist.all_gather(parts, out_local.contiguous(), group=group)
    return parts[0], parts[1]     # group rank 0 = even global rank = conditional branch
