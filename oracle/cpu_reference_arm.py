"""CPU arm of bench.py: the reference's SIMULATED-quant denoising path timed on the host cores.  TEST / BENCH
INFRASTRUCTURE ONLY (bench.py's `cpu_baseline` leg and `--impl reference`); nothing under vidit-q_b200/ imports it.

The Python reference cannot travel to the GPU box (/root/reference does not exist there), so this is a PORT
(`cpu_baseline.kind = "port"`): the reference's graph (STDiT-XL/2 blocks at 16x512x512: 16384 tokens, hidden 1152; the
state_dict-compatible restatement in vidit-q_b200/stdit.py run on CPU in fp32, un-fused `forward` schedule = the reference's
13 linears per block, stdit.py:96-133) with every quantised linear executing oracle.torch_fake_quant.quant_linear_fake —
the reference's QuantLayer arithmetic op for op (per-token dynamic activation fake-quant with batch-pooled statistics,
per-channel weight fake-quant RE-EXECUTED EVERY CALL as quant_layer.py:185 does, then F.linear), pinned bit-exact to the
unmodified reference classes by tests/test_oracle_golden.py.  fp32 and all host threads, as BASELINE.md section 4 plans
(CPU fp16 GEMM is not representative).

One bounded sample = embed + `n_blocks` full-size blocks + final layer of ONE forward; a denoise step (cfg_split: two
forwards of 28 blocks) is extrapolated linearly from the per-block time, and the line says so.
"""
import os
import time

import torch

T_FRAMES, S_TOKENS, HIDDEN, DEPTH, PROMPT_LEN = 16, 1024, 1152, 28, 120
FP_LAYERS = ["x_embedder", "t_block", "t_embedder", "y_embedder", "final_layer"]

_state = {}


class _Cfg(dict):
    __getattr__ = dict.get


def _build(n_blocks):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import torch_fake_quant as TF
    from viditq_b200.qdiff import QuantModel
    from viditq_b200.stdit import STDiT
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count() or 1)
    model = STDiT(input_size=(T_FRAMES, 64, 64), depth=n_blocks, hidden_size=HIDDEN, num_heads=16)
    model.eval()
    sq = _Cfg(enable=False, channel_wise_scale_type="momentum_act_max", momentum=0.95, alpha=0.625)
    wq = _Cfg(n_bits=8, per_group="channel", channel_dim=0, scale_method="min_max", round_mode="nearest",
              mixed_precision=[4, 6, 8])
    aq = _Cfg(n_bits=8, per_group="token", scale_method="min_max", round_mode="nearest_ste", running_stat=False,
              dynamic=True, sym=False, n_spatial_token=S_TOKENS, n_temporal_token=T_FRAMES, n_prompt=PROMPT_LEN,
              smooth_quant=sq)
    qnn = QuantModel(model, wq, aq)
    qnn.set_module_name_for_quantizer(module=qnn.model)
    qnn.init_weight_quant_params()
    qnn.set_quant_init_done("weight")
    qnn.set_quant_init_done("activation")
    qnn.set_quant_state(True, True)
    qnn.set_layer_quant(model=qnn, module_name_list=FP_LAYERS, quant_level="per_layer", weight_quant=False,
                        act_quant=False, prefix="")

    def make(layer, plain):
        def fwd(inp, *a, **k):
            if not (layer.weight_quant and layer.act_quant):
                return plain(inp)
            G, rows = layer._pool_view(inp)
            wq_ = layer.weight_quantizer
            out = TF.quant_linear_fake(inp.reshape(G, rows, inp.shape[-1]), layer.weight, layer.bias, wq_.delta,
                                       wq_.zero_point, wq_.n_bits, layer.act_quantizer.n_bits)
            return out.reshape(*inp.shape[:-1], -1)
        return fwd
    for _, layer in qnn.quant_layers():
        layer.forward = make(layer, layer.forward)
    times = {"blocks": 0.0}

    def pre(mod, args):
        mod._t0 = time.perf_counter()

    def post(mod, args, out):
        times["blocks"] += time.perf_counter() - mod._t0
    for blk in model.blocks:
        blk.register_forward_pre_hook(pre)
        blk.register_forward_hook(post)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, T_FRAMES, 64, 64, generator=g)
    y = torch.randn(1, 1, PROMPT_LEN, 4096, generator=g)
    mask = torch.zeros(1, PROMPT_LEN, dtype=torch.int64)
    mask[0, :109] = 1
    return qnn, times, (x, torch.tensor([500.0]), y, mask)


def time_sample(n_blocks=1):
    """Run one bounded sample; returns dict(seconds, block_seconds, rest_seconds, n_blocks, cores, tokens)."""
    if n_blocks not in _state:
        _state[n_blocks] = _build(n_blocks)
    qnn, times, (x, t, y, mask) = _state[n_blocks]
    times["blocks"] = 0.0
    t0 = time.perf_counter()
    qnn(x, t, y, mask=mask)
    total = time.perf_counter() - t0
    return {"seconds": total, "block_seconds": times["blocks"] / n_blocks, "rest_seconds": total - times["blocks"],
            "n_blocks": n_blocks, "cores": torch.get_num_threads(), "tokens": T_FRAMES * S_TOKENS}


def steps_per_sec(sample):
    """Denoise steps per second extrapolated from a sample: one step = 2 forwards x (embed + final + 28 blocks)."""
    return 1.0 / (2.0 * (sample["rest_seconds"] + DEPTH * sample["block_seconds"]))


def describe(sample):
    return (f"torch fp32 port of the reference's simulated-quant STDiT forward, {sample['cores']} threads: embed + "
            f"{sample['n_blocks']} full-size block(s) ({sample['tokens']} tokens) + final layer in {sample['seconds']:.2f} s "
            f"({sample['block_seconds']:.2f} s per block); step = 2 forwards x (rest + 28 blocks), extrapolated linearly")
