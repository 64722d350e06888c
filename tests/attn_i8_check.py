"""TEST INFRASTRUCTURE (lives under tests/ because it executes the oracle): helper of tests/test_gpu_attn_i8.py and
bring-up / timing script of the opt-in INT8 attention (vq_attn_spatial_i8) — operand codes against
oracle/attn_i8_oracle.py, the attention kernel against the oracle's restatement of its arithmetic on the kernel's own
codes, the scheme against fp16 attention, and the launch times beside vq_attn_spatial.
    python tests/attn_i8_check.py [--full] [--no-timing]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import attn_i8_oracle as A   # noqa: E402
from viditq_b200 import ops              # noqa: E402


def read_workspace(ws, n_seq, S, H, D=72):
    off = ops.attn_i8_workspace_layout(n_seq, S, H, D)
    rows, C = n_seq * S, H * D
    b = ws.cpu()
    qk = b[off["qk8"]:off["qk8"] + rows * 2 * H * 80].view(torch.int8).reshape(rows, 2, H, 80)
    vt = b[off["vt8"]:off["vt8"] + n_seq * C * S].view(torch.int8).reshape(n_seq, H, D, S)
    f = lambda name, n: b[off[name]:off[name] + 4 * n].view(torch.float32)
    return dict(q8=qk[:, 0, :, :D].reshape(n_seq, S, H, D).float(), k8=qk[:, 1, :, :D].reshape(n_seq, S, H, D).float(),
                pad=qk[..., D:], v8=vt.permute(0, 3, 1, 2).float(), sq=f("sq", rows * H).reshape(n_seq, S, H),
                sk=f("sk", rows // 64 * H).reshape(n_seq, S // 64, H), sv=f("sv", n_seq * C).reshape(n_seq, H, D),
                kmean=f("kmean", n_seq * C).reshape(n_seq, H, D))


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item(), ((a - b).abs().max() / b.abs().max()).item()


def check(n_seq, S, H, seed, gain=1.0, verbose=True):
    D = 72
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n_seq * S, 3 * H * D, generator=g) * gain)
    x[:, ::7] *= 3.0                          # channel outliers
    x[:, H * D:2 * H * D] += 1.5              # a common K offset: what the mean subtraction removes
    x = x.half()
    xd = x.cuda()
    nbytes = ops._lib.lib().vq_attn_i8_workspace_bytes(n_seq, S, H, D)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    out = ops.attn_spatial_i8(xd, n_seq, S, H, D, D ** -0.5, workspace=ws)
    torch.cuda.synchronize()
    z = read_workspace(ws, n_seq, S, H)
    ref_mean = x.float().reshape(n_seq, S, 3, H, D)[:, :, 1].mean(dim=1)
    zo = A.quantise_qkv(x, n_seq, S, H, D, kmean=z["kmean"])
    res = {"kmean_err": (z["kmean"] - ref_mean).abs().max().item(), "pad_zero": bool((z["pad"] == 0).all())}
    for n in ("q8", "k8", "v8", "sq", "sk", "sv"):
        res[n + "_mismatch"] = int((z[n] != zo[n]).sum())
    tiled = A.attention_i8_tiled(z, D ** -0.5)
    o = out.float().cpu()
    res["kernel_vs_tiled_oracle"] = rel(o, tiled)
    res["kernel_vs_fp"] = rel(o, A.attention_fp(x, n_seq, S, H, D ** -0.5))
    res["scheme_vs_fp"] = rel(A.attention_i8(x, n_seq, S, H, D ** -0.5), A.attention_fp(x, n_seq, S, H, D ** -0.5))
    f16 = ops.attn_spatial(xd, n_seq, S, H, D, D ** -0.5).float().cpu()
    res["fp16_kernel_vs_fp"] = rel(f16, A.attention_fp(x, n_seq, S, H, D ** -0.5))
    if verbose:
        print(f"n_seq={n_seq} S={S} H={H}:", res, flush=True)
        if res["kernel_vs_tiled_oracle"][0] > 5e-3:
            d = (o - tiled).reshape(n_seq, S, H, D)
            t = tiled.reshape(n_seq, S, H, D)
            print("  per head rel-L2:", [round((d[:, :, h].norm() / t[:, :, h].norm()).item(), 4) for h in range(H)])
            print("  per 128-row tile:", [round((d[:, r:r + 128].norm() / t[:, r:r + 128].norm()).item(), 4)
                                          for r in range(0, S, 128)])
            print("  per dim block of 8:", [round((d[..., c:c + 8].norm() / t[..., c:c + 8].norm()).item(), 4)
                                            for c in range(0, D, 8)])
            print("  finite:", bool(torch.isfinite(o).all()), " |o| max", o.abs().max().item(), " |ref| max",
                  tiled.abs().max().item())
    return res


def timing(n_seq=32, S=1024, H=16, iters=20):
    D = 72
    x = (torch.randn(n_seq * S, 3 * H * D, device="cuda")).half()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    L = ops._lib.lib()
    nbytes = L.vq_attn_i8_workspace_bytes(n_seq, S, H, D)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(n_seq * S, H * D, dtype=torch.float16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def t(fn):
        for _ in range(3):
            fn()
        ms = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        ms.sort()
        return 1e3 * ms[len(ms) // 2]
    res = {
        "fp16 vq_attn_spatial us": t(lambda: ops.attn_spatial(x, n_seq, S, H, D, D ** -0.5, out=out)),
        "i8 quantise (2 launches) us": t(lambda: L.vq_attn_i8_quantise(x.data_ptr(), ws.data_ptr(), n_seq, S, H, D, st)),
        "i8 attend us": t(lambda: L.vq_attn_i8_attend(ws.data_ptr(), out.data_ptr(), n_seq, S, H, D, D ** -0.5, st)),
        "i8 total us": t(lambda: ops.attn_spatial_i8(x, n_seq, S, H, D, D ** -0.5, out=out, workspace=ws)),
    }
    print(f"timing n_seq={n_seq} S={S} H={H} (L2 flushed between launches):", res, flush=True)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--no-timing", action="store_true")
    a = ap.parse_args()
    check(1, 256, 1, 0)
    check(1, 512, 2, 1)
    check(2, 1024, 4, 2, gain=2.0)
    if a.full:
        check(3, 1024, 16, 3)
    if not a.no_timing:
        timing()
