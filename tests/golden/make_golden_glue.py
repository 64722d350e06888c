"""Generate tests/golden/glue_golden.npz: outputs of the UNMODIFIED reference Attention (eager branch) and PatchEmbed3D
classes (t2v/opensora/models/layers/blocks.py) on seeded inputs, fp32 on CPU.  The q / k / v the attention core sees and
the tensor it hands to `proj` are captured with forward hooks.  Run here:  python tests/golden/make_golden_glue.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install_opensora()
from opensora.models.layers.blocks import Attention, PatchEmbed3D  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "glue_golden.npz")


def main():
    torch.set_grad_enabled(False)
    torch.manual_seed(2024)
    C, H = 1152, 16
    attn = Attention(C, num_heads=H, qkv_bias=True, enable_flashattn=False).eval()
    for lin in (attn.q, attn.k, attn.v):
        lin.weight.mul_(3.0)                      # spread the logits
    cap = {}
    for name in ("q", "k", "v"):
        getattr(attn, name).register_forward_hook(lambda m, i, o, n=name: cap.__setitem__(n, o.clone()))
    attn.proj.register_forward_hook(lambda m, i, o: cap.__setitem__("core", i[0].clone()))
    x = torch.randn(1, 80, C)
    attn(x)
    pe = PatchEmbed3D(patch_size=(1, 2, 2), in_chans=4, embed_dim=C).eval()
    pe.proj.weight.copy_(pe.proj.weight.half().float())      # fp16-representable parameters (the model runs in half)
    pe.proj.bias.copy_(pe.proj.bias.half().float())
    z = torch.randn(1, 4, 3, 8, 12).half().float()
    y = pe(z)
    np.savez_compressed(OUT, q=cap["q"].numpy(), k=cap["k"].numpy(), v=cap["v"].numpy(), core=cap["core"].numpy(),
                        heads=np.int64(H), z=z.numpy(), pe_w=pe.proj.weight.numpy(), pe_b=pe.proj.bias.numpy(),
                        pe_out=y.numpy())
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
