#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``

It imports the UNMODIFIED ``/root/reference/src/cara/cara.py`` -- the reference's
own ``cara`` / ``set_cara`` / ``cp_attn`` / ``cp_mlp`` bytes -- with the two
un-vendored pip dependencies it needs (timm 0.4.12, tensorly 0.8.1) provided by
``oracle/shims``.  Inputs are the deterministic numpy-PCG64 tensors of
``oracle.cara_oracle.synthetic_state`` / ``synthetic_batch`` (inputs only: every
OUTPUT stored here is computed by the reference's code, in ``.eval()`` mode so
weight-dropout and DropPath are identities, with autograd for the gradients as in
vit_cp.py:47-49).

Files written (float64/float32 ``.npz``, a few hundred KB each):
  ref_vitb_d2_r8_fp64.npz    ViT-B width, depth 2, rank 8, 10 classes, B=2, fp64
  ref_vitb_d12_r16_fp32.npz  full ViT-B/16, rank 16, 100 classes, B=2, fp32
  ref_halves_fp64.npz        cp_attn / cp_mlp bound forwards of block 1 on x[2,5,768]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
import importlib.util  # noqa: E402

REF_CARA = "/root/reference/src/cara/cara.py"
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, ROOT)

from timm.models import create_model  # noqa: E402  (the shim)

# The repo's own ``src`` package shadows the reference's ``src`` namespace package on sys.path, so the unmodified
# reference module is loaded by FILE PATH (under a private module name; it only imports torch / timm / tensorly).
_spec = importlib.util.spec_from_file_location("_reference_cara", REF_CARA)
ref_cara = importlib.util.module_from_spec(_spec)
sys.modules["_reference_cara"] = ref_cara
_spec.loader.exec_module(ref_cara)

from oracle import cara_oracle as O  # noqa: E402  (inputs only)

assert os.path.realpath(ref_cara.__file__) == REF_CARA, ref_cara.__file__


def build_reference(g: O.Geometry, scale: float, dtype):
    vit = create_model("vit_base_patch16_224_in21k", drop_path_rate=0.1, depth=g.depth)
    vit = ref_cara.cara({"model": vit, "rank": g.rank, "scale": scale, "l_mu": 1.0, "l_std": 0.0})
    vit.reset_classifier(g.num_classes)
    st = O.synthetic_state(g, dtype=dtype)
    # the reference declares A1[36,R] / P1[108,R] whatever the depth (cara.py:112,116)
    full = vit.state_dict()
    for k, v in st.items():
        if k in ("CP_A1", "CP_P1"):
            buf = torch.zeros_like(full[k], dtype=dtype)
            buf[: v.shape[0]] = v
            st[k] = buf
    vit = vit.to(dtype)
    missing, unexpected = vit.load_state_dict(st, strict=True), None
    vit.eval()
    return vit, st


def run_model(g, scale, dtype, batch, out):
    vit, st = build_reference(g, scale, dtype)
    x, y = O.synthetic_batch(g, batch, dtype=dtype)
    trainable = {}
    for n, p in vit.named_parameters():          # vit_cp.py:176-182
        if "CP" in n or "head" in n:
            trainable[n] = p
        else:
            p.requires_grad = False
    logits = vit(x)                               # vit_cp.py:46
    loss = torch.nn.functional.cross_entropy(logits, y)   # :47
    loss.backward()                               # :49
    rec = {"logits": logits.detach().numpy(), "loss": loss.detach().numpy(),
           "scale": np.float64(scale), "batch": np.int64(batch)}
    L = g.depth
    for n, p in trainable.items():
        gr = p.grad.detach()
        if n == "CP_A1":
            gr = gr[: 3 * L]
        if n == "CP_P1":
            gr = gr[: 9 * L]
        rec["grad." + n] = gr.numpy()
    np.savez_compressed(os.path.join(HERE, out), **rec)
    print(out, "loss", float(loss), "logits[0,:4]", logits[0, :4].tolist())


def run_halves(out):
    g = O.Geometry(depth=2, rank=8, num_classes=10)
    vit, st = build_reference(g, 2.5, torch.float64)
    rng = np.random.Generator(np.random.PCG64(99))
    x = torch.from_numpy(rng.standard_normal((2, 5, 768)))
    blk = vit.blocks[1]
    with torch.no_grad():
        a = blk.attn(x)      # bound cp_attn (cara.py:155-156)
        m = blk.mlp(x)       # bound cp_mlp (cara.py:163-164)
    np.savez_compressed(os.path.join(HERE, out), x=x.numpy(), attn=a.numpy(), mlp=m.numpy(),
                        scale=np.float64(2.5), layer=np.int64(1),
                        attn_idx=np.int64(blk.attn.attn_idx), idx=np.int64(blk.attn.idx),
                        mlp_idx=np.int64(blk.mlp.idx))
    print(out, "attn_idx", blk.attn.attn_idx, "idx", blk.attn.idx, "mlp.idx", blk.mlp.idx)


if __name__ == "__main__":
    if len(sys.argv) > 1:          # write somewhere else (e.g. to compare against the committed files)
        HERE = os.path.abspath(sys.argv[1])
        os.makedirs(HERE, exist_ok=True)
    torch.manual_seed(0)
    run_model(O.Geometry(depth=2, rank=8, num_classes=10), 2.5, torch.float64, 2, "ref_vitb_d2_r8_fp64.npz")
    run_model(O.Geometry(depth=12, rank=16, num_classes=100), 1.0, torch.float32, 2, "ref_vitb_d12_r16_fp32.npz")
    run_halves("ref_halves_fp64.npz")
