"""CPU oracle for the CaRA fine-tuning hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only
as the checker / the CPU baseline -- never as the product path.  The product
(``cara_b200``) fails loudly when its CUDA library is missing; it never falls
back to this file.

What it restates (plain PyTorch on the CPU, fp32 or fp64, autograd for the
backward), following the reference line by line in *behaviour* but written
independently and generalised from the reference's hard-coded ViT-B shapes
(cara.py:112-125) to any ``C = H*D`` / ``hidden = 4C`` geometry:

* ``cp_to_tensor``           tensorly 0.8.1 ``cp_tensor.py`` (un-vendored pip dep,
                             pyproject.toml:11; call sites cara.py:27,52,76,88)
* ``attn_half`` / ``mlp_half``  cara.py:15-60 (``cp_attn``) and cara.py:63-95 (``cp_mlp``):
                             the delta tensors are MATERIALISED exactly as the
                             reference does (Khatri-Rao reconstruction, optional
                             element dropout on the delta *weights*, dense einsum)
                             -- deliberately not the factored chain the CUDA
                             kernels use, so the two are independent statements.
* ``declare_cp`` / row maps  cara.py:98-166 (``set_cara``): parameter shapes, init,
                             ``attn_idx = 3l``, ``idx = 9l``, ``mlp.idx = 9l+1``
* backbone                   timm 0.4.12 ``VisionTransformer``/``Block``/``Attention``/
                             ``Mlp``/``DropPath`` (un-vendored pip dep, pyproject.toml:12)
* ``train_step``             image_classification/vit_cp.py:45-50 and :176-185
                             (CE loss, grads only for ``CP*`` + ``head``, AdamW wd 1e-4)

Parity pinning: the reference's own tests (tests/test_cara.py) hold no numeric
vectors, so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: the
unmodified /root/reference/src/cara/cara.py is imported (with the
``oracle/shims`` stand-ins for timm/tensorly) by ``tests/golden/make_golden.py``
and its forward/backward results are committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- geometry


@dataclass(frozen=True)
class Geometry:
    """Shapes of one ViT + CaRA instance."""

    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    patch: int = 16
    img: int = 224
    num_classes: int = 100
    rank: int = 16
    in_chans: int = 3

    @property
    def tokens(self) -> int:
        return (self.img // self.patch) ** 2 + 1

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def hidden(self) -> int:
        return 4 * self.embed_dim


VIT_B16 = dict(embed_dim=768, depth=12, num_heads=12, patch=16)
VIT_L16 = dict(embed_dim=1024, depth=24, num_heads=16, patch=16)
VIT_H14 = dict(embed_dim=1280, depth=32, num_heads=16, patch=14)

CP_NAMES = ("CP_A1", "CP_A2", "CP_A3", "CP_A4", "CP_P1", "CP_P2", "CP_P3",
            "CP_R1", "CP_R2", "CP_bias1", "CP_bias2", "CP_bias3")


def cp_shapes(g: Geometry) -> Dict[str, Tuple[int, ...]]:
    """cara.py:112-125 with 36 -> 3L, 108 -> 9L, 768 -> C, 12 -> H, 64 -> D."""
    C, L, H, D, R = g.embed_dim, g.depth, g.num_heads, g.head_dim, g.rank
    return {
        "CP_A1": (3 * L, R), "CP_A2": (C, R), "CP_A3": (H, R), "CP_A4": (D, R),
        "CP_P1": (9 * L, R), "CP_P2": (C, R), "CP_P3": (C, R),
        "CP_R1": (R,), "CP_R2": (R,),
        "CP_bias1": (C,), "CP_bias2": (4 * C,), "CP_bias3": (C,),
    }


def backbone_shapes(g: Geometry) -> Dict[str, Tuple[int, ...]]:
    """timm 0.4.12 state_dict key schema (164 tensors for ViT-B incl. CP_*)."""
    C, P = g.embed_dim, g.patch
    s: Dict[str, Tuple[int, ...]] = {
        "cls_token": (1, 1, C), "pos_embed": (1, g.tokens, C),
        "patch_embed.proj.weight": (C, g.in_chans, P, P), "patch_embed.proj.bias": (C,),
    }
    for i in range(g.depth):
        b = "blocks.%d." % i
        s[b + "norm1.weight"] = (C,); s[b + "norm1.bias"] = (C,)
        s[b + "attn.qkv.weight"] = (3 * C, C); s[b + "attn.qkv.bias"] = (3 * C,)
        s[b + "attn.proj.weight"] = (C, C); s[b + "attn.proj.bias"] = (C,)
        s[b + "norm2.weight"] = (C,); s[b + "norm2.bias"] = (C,)
        s[b + "mlp.fc1.weight"] = (4 * C, C); s[b + "mlp.fc1.bias"] = (4 * C,)
        s[b + "mlp.fc2.weight"] = (C, 4 * C); s[b + "mlp.fc2.bias"] = (C,)
    s["norm.weight"] = (C,); s["norm.bias"] = (C,)
    s["head.weight"] = (g.num_classes, C); s["head.bias"] = (g.num_classes,)
    return s


# --------------------------------------------------------------------------- synthetic state


def _normal(rng: np.random.Generator, shape, std=1.0, mean=0.0) -> Tensor:
    return torch.from_numpy(rng.standard_normal(shape) * std + mean)


def synthetic_state(g: Geometry, seed: int = 0, cp_seed: int = 1234,
                    dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic (numpy PCG64) random-init weights + NON-default CP factors.

    Backbone: timm-style scales (Linear std .02, pos/cls std .02) plus randomised
    biases / LayerNorm affine so no term is trivially absent.
    CP factors: the D.3 recipe (delta-weight entry std ~0.01); the reference's own
    default init (cara.py:127-142) makes every delta exactly zero, which would not
    exercise the adapter at all.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    st: Dict[str, Tensor] = {}
    for k, shp in backbone_shapes(g).items():
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
            st[k] = _normal(rng, shp, 0.1, 1.0)
        elif k.endswith(".bias"):
            st[k] = _normal(rng, shp, 0.02)
        elif k == "patch_embed.proj.weight":
            st[k] = _normal(rng, shp, 1.0 / math.sqrt(g.in_chans * g.patch * g.patch))
        else:
            st[k] = _normal(rng, shp, 0.02)
    rng = np.random.Generator(np.random.PCG64(cp_seed))
    R = g.rank
    sa = (0.01 / math.sqrt(R)) ** 0.25
    sp = (0.01 / math.sqrt(R)) ** (1.0 / 3.0)
    for k, shp in cp_shapes(g).items():
        if k in ("CP_R1", "CP_R2"):
            st[k] = _normal(rng, shp, 0.1, 1.0)
        elif k.startswith("CP_bias"):
            st[k] = _normal(rng, shp, 0.02)
        elif k.startswith("CP_A"):
            st[k] = _normal(rng, shp, sa)
        else:
            st[k] = _normal(rng, shp, sp)
    return {k: v.to(dtype).contiguous() for k, v in st.items()}


def synthetic_batch(g: Geometry, batch: int, seed: int = 4321,
                    dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """x ~ N(0,1) [B,3,img,img]; labels uniform in [0, classes)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = torch.from_numpy(rng.standard_normal((batch, g.in_chans, g.img, g.img))).to(dtype)
    y = torch.from_numpy(rng.integers(0, g.num_classes, size=(batch,), dtype=np.int64))
    return x, y


def declare_cp(g: Geometry, l_mu: float = 1.0, l_std: float = 0.0,
               generator: Optional[torch.Generator] = None,
               dtype=torch.float32) -> Dict[str, Tensor]:
    """Default CaRA init, cara.py:127-142 (xavier-normal A1/P1, zero A2/P2,
    orthogonal A3/A4/P3, lambda ~ N(mu, std) or ones, zero biases)."""
    out: Dict[str, Tensor] = {}
    for k, shp in cp_shapes(g).items():
        t = torch.empty(shp, dtype=torch.float32)
        if k in ("CP_A1", "CP_P1"):
            torch.nn.init.xavier_normal_(t, generator=generator)
        elif k in ("CP_A2", "CP_P2") or k.startswith("CP_bias"):
            t.zero_()
        elif k in ("CP_A3", "CP_A4", "CP_P3"):
            torch.nn.init.orthogonal_(t, generator=generator)
        else:  # CP_R1 / CP_R2, cara.py:134-139
            if l_std != 0.0:
                torch.nn.init.normal_(t, mean=l_mu, std=l_std, generator=generator)
            elif l_mu == 1.0:
                t.fill_(1.0)
            # else: the reference leaves the memory uninitialised (cara.py:134-139 quirk)
        out[k] = t.to(dtype)
    return out


# --------------------------------------------------------------------------- CP reconstruction


def khatri_rao(mats) -> Tensor:
    """Column-wise Kronecker product; row index in C order, first matrix slowest."""
    out = mats[0]
    for m in mats[1:]:
        out = torch.einsum("ir,jr->ijr", out, m).reshape(-1, out.shape[1])
    return out


def cp_to_tensor(weights: Tensor, factors) -> Tensor:
    """tensorly 0.8.1 ``cp_to_tensor``: sum_r w_r * f0[:,r] o f1[:,r] o ... (dense)."""
    ranks = {f.shape[1] for f in factors}
    if len(ranks) != 1:
        raise ValueError("All the factors of a CP tensor should have the same number of column")
    flat = (factors[0] * weights) @ khatri_rao(list(factors[1:])).T
    return flat.reshape([f.shape[0] for f in factors])


def _wdrop(t: Tensor, p: float, train: bool) -> Tensor:
    """``self.dp`` = nn.Dropout(0.1) applied to the materialised delta (cara.py:35,57,81,92)."""
    return F.dropout(t, p, training=train) if (train and p > 0.0) else t


# --------------------------------------------------------------------------- the two adapted halves


def attn_half(st: Dict[str, Tensor], g: Geometry, layer: int, x: Tensor, scale: float,
              train: bool = False, wdrop: float = 0.0) -> Tensor:
    """cara.py:24-59.  x: [B,N,C] (already LayerNorm'ed) -> attention branch output."""
    B, N, C = x.shape
    H, D = g.num_heads, g.head_dim
    b = "blocks.%d.attn." % layer
    qkv = F.linear(x, st[b + "qkv.weight"], st[b + "qkv.bias"])                 # :25
    a1 = st["CP_A1"][3 * layer:3 * layer + 3]                                      # :26 attn_idx = 3l
    dW = cp_to_tensor(st["CP_R1"], (a1, st["CP_A2"], st["CP_A3"], st["CP_A4"]))   # :27-32 [3,C,H,D]
    dW = dW.reshape(3, C, H * D)                                                   # :33-34
    delta = torch.einsum("bnd,kde->kbne", x, _wdrop(dW, wdrop, train))             # :35
    delta = delta.reshape(3, B, N, H, D).permute(0, 1, 3, 2, 4)                    # :36-38
    qkv = qkv.reshape(B, N, 3, H, D).permute(2, 0, 3, 1, 4)                        # :39-41
    qkv = qkv + delta * scale                                                      # :42
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = ((q @ k.transpose(-2, -1)) * (D ** -0.5)).softmax(dim=-1)                # :44-46 (attn_drop p=0)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)                                 # :48
    proj = F.linear(o, st[b + "proj.weight"], st[b + "proj.bias"])                # :50
    p1 = st["CP_P1"][9 * layer:9 * layer + 1]                                      # :51 idx = 9l
    dP = cp_to_tensor(st["CP_R2"], (p1, st["CP_P2"], st["CP_P3"])).reshape(C, C)   # :52-56
    pd = o @ _wdrop(dP.T, wdrop, train) + st["CP_bias1"]                           # :57
    return proj + pd * scale                                                       # :58 (proj_drop p=0)


def mlp_half(st: Dict[str, Tensor], g: Geometry, layer: int, x: Tensor, scale: float,
             train: bool = False, wdrop: float = 0.0) -> Tensor:
    """cara.py:72-95.  x: [B,N,C] (already LayerNorm'ed) -> FFN branch output."""
    C = g.embed_dim
    b = "blocks.%d.mlp." % layer
    i0 = 9 * layer + 1                                                             # mlp.idx = 9l+1
    p_up, p_dn = st["CP_P1"][i0:i0 + 4], st["CP_P1"][i0 + 4:i0 + 8]                # :72-73
    up = F.linear(x, st[b + "fc1.weight"], st[b + "fc1.bias"])                    # :75
    dU = cp_to_tensor(st["CP_R2"], (p_up, st["CP_P2"], st["CP_P3"])).reshape(4 * C, C)   # :76-80
    up = up + (x @ _wdrop(dU.T, wdrop, train) + st["CP_bias2"]) * scale            # :81-82
    h = F.gelu(up)                                                                 # :84 exact erf
    dn = F.linear(h, st[b + "fc2.weight"], st[b + "fc2.bias"])                    # :87
    dD = cp_to_tensor(st["CP_R2"], (p_dn, st["CP_P2"], st["CP_P3"])).reshape(4 * C, C)   # :88-91
    return dn + (h @ _wdrop(dD, wdrop, train) + st["CP_bias3"]) * scale            # :92-93


# --------------------------------------------------------------------------- backbone


def patch_embed(st: Dict[str, Tensor], g: Geometry, img: Tensor) -> Tensor:
    """timm PatchEmbed + cls token + position embedding -> [B,N,C]."""
    x = F.conv2d(img, st["patch_embed.proj.weight"], st["patch_embed.proj.bias"], stride=g.patch)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((st["cls_token"].expand(x.shape[0], -1, -1), x), dim=1)
    return x + st["pos_embed"]


def drop_path_rates(g: Geometry, rate: float):
    return [float(r) for r in torch.linspace(0, rate, g.depth)]


def forward(st: Dict[str, Tensor], g: Geometry, img: Tensor, scale: float = 1.0, *,
            train: bool = False, wdrop: float = 0.0, drop_path: float = 0.0,
            keep: Optional[Tensor] = None) -> Tensor:
    """Logits.  ``train`` + ``wdrop``/``drop_path`` reproduce the stochastic train mode
    (vit_cp.py:155 rate 0.1, cara.py:148 p 0.1); parity runs use the deterministic
    path.  ``keep``: optional explicit DropPath multipliers [depth, 2, B] (already
    divided by keep-prob) so a stochastic-depth run can be replayed exactly."""
    x = patch_embed(st, g, img)
    rates = drop_path_rates(g, drop_path)
    for l in range(g.depth):
        b = "blocks.%d." % l
        for j, (norm, half) in enumerate((("norm1", attn_half), ("norm2", mlp_half))):
            h = F.layer_norm(x, (g.embed_dim,), st[b + norm + ".weight"], st[b + norm + ".bias"], 1e-6)
            y = half(st, g, l, h, scale, train, wdrop)
            if keep is not None:
                y = y * keep[l, j].to(y.dtype).view(-1, 1, 1)
            elif train and rates[l] > 0.0:
                kp = 1.0 - rates[l]
                y = y / kp * torch.floor(kp + torch.rand(y.shape[0], 1, 1, dtype=y.dtype))
            x = x + y
    x = F.layer_norm(x, (g.embed_dim,), st["norm.weight"], st["norm.bias"], 1e-6)
    return F.linear(x[:, 0], st["head.weight"], st["head.bias"])


def trainable_names(st: Dict[str, Tensor]):
    """vit_cp.py:176-182: trainable iff the name contains "CP" or "head"."""
    return [k for k in st if ("CP" in k or "head" in k)]


def loss_and_grads(st: Dict[str, Tensor], g: Geometry, img: Tensor, labels: Tensor,
                   scale: float = 1.0, **fw) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """vit_cp.py:46-49: logits, mean cross-entropy, autograd grads for CP* + head."""
    names = trainable_names(st)
    leaves = {k: st[k].detach().clone().requires_grad_(True) for k in names}
    work = dict(st); work.update(leaves)
    logits = forward(work, g, img, scale, **fw)
    loss = F.cross_entropy(logits, labels)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    return logits.detach(), loss.detach(), dict(zip(names, grads))


def adamw_update(p: Tensor, grad: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-3,
                 wd: float = 1e-4, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.AdamW single-tensor update (vit_cp.py:185, default betas/eps)."""
    p = p * (1.0 - lr * wd)
    m = b1 * m + (1.0 - b1) * grad
    v = b2 * v + (1.0 - b2) * grad * grad
    denom = (v.sqrt() / math.sqrt(1.0 - b2 ** step)) + eps
    p = p - (lr / (1.0 - b1 ** step)) * m / denom
    return p, m, v


def train_step(st: Dict[str, Tensor], opt: Dict[str, Tuple[Tensor, Tensor]], step: int,
               g: Geometry, img: Tensor, labels: Tensor, scale: float = 1.0,
               lr: float = 1e-3, wd: float = 1e-4, **fw):
    """One vit_cp.py:45-50 iteration; returns (loss, logits) and updates st/opt in place."""
    logits, loss, grads = loss_and_grads(st, g, img, labels, scale, **fw)
    for k, gr in grads.items():
        m, v = opt.get(k, (torch.zeros_like(st[k]), torch.zeros_like(st[k])))
        st[k], m, v = adamw_update(st[k], gr, m, v, step, lr, wd)
        opt[k] = (m, v)
    return loss, logits


# --------------------------------------------------------------------------- factored view (A.1-A.3)


def adapter_terms(st: Dict[str, Tensor], g: Geometry, layer: int, which: str):
    """The per-projection (A [K,R], c [slices,R], B [N/slices,R], beta) of SURVEY A.1,
    derived from the SAME index arithmetic as attn_half/mlp_half; used by tests to
    check the C-ABI's factor staging and by ``merged_weights``."""
    C = g.embed_dim
    if which == "qkv":
        return (st["CP_A2"], st["CP_R1"] * st["CP_A1"][3 * layer:3 * layer + 3],
                khatri_rao([st["CP_A3"], st["CP_A4"]]), None)
    if which == "proj":
        return (st["CP_P3"], st["CP_R2"] * st["CP_P1"][9 * layer:9 * layer + 1], st["CP_P2"], st["CP_bias1"])
    if which == "fc1":
        return (st["CP_P3"], st["CP_R2"] * st["CP_P1"][9 * layer + 1:9 * layer + 5], st["CP_P2"], st["CP_bias2"])
    if which == "fc2":
        a = khatri_rao([st["CP_P1"][9 * layer + 5:9 * layer + 9], st["CP_P2"]])
        return (a, st["CP_R2"].unsqueeze(0), st["CP_P3"], st["CP_bias3"])
    raise KeyError(which)


def merged_weights(st: Dict[str, Tensor], g: Geometry, scale: float) -> Dict[str, Tensor]:
    """Eval-mode fold W_eff = W + s*dW, b_eff = b + s*beta built from the
    MATERIALISED deltas (same cp_to_tensor calls as the two halves)."""
    C, H, D = g.embed_dim, g.num_heads, g.head_dim
    out = {k: v.clone() for k, v in st.items() if not k.startswith("CP_")}
    for l in range(g.depth):
        b = "blocks.%d." % l
        dW = cp_to_tensor(st["CP_R1"], (st["CP_A1"][3 * l:3 * l + 3], st["CP_A2"], st["CP_A3"], st["CP_A4"]))
        out[b + "attn.qkv.weight"] += scale * dW.reshape(3, C, H * D).permute(0, 2, 1).reshape(3 * C, C)
        dP = cp_to_tensor(st["CP_R2"], (st["CP_P1"][9 * l:9 * l + 1], st["CP_P2"], st["CP_P3"])).reshape(C, C)
        out[b + "attn.proj.weight"] += scale * dP
        out[b + "attn.proj.bias"] += scale * st["CP_bias1"]
        dU = cp_to_tensor(st["CP_R2"], (st["CP_P1"][9 * l + 1:9 * l + 5], st["CP_P2"], st["CP_P3"])).reshape(4 * C, C)
        out[b + "mlp.fc1.weight"] += scale * dU
        out[b + "mlp.fc1.bias"] += scale * st["CP_bias2"]
        dD = cp_to_tensor(st["CP_R2"], (st["CP_P1"][9 * l + 5:9 * l + 9], st["CP_P2"], st["CP_P3"])).reshape(4 * C, C)
        out[b + "mlp.fc2.weight"] += scale * dD.T
        out[b + "mlp.fc2.bias"] += scale * st["CP_bias3"]
    return out


def forward_plain(st: Dict[str, Tensor], g: Geometry, img: Tensor) -> Tensor:
    """Un-adapted timm ViT forward over a (possibly merged) backbone state."""
    C, H, D = g.embed_dim, g.num_heads, g.head_dim
    x = patch_embed(st, g, img)
    B, N, _ = x.shape
    for l in range(g.depth):
        b = "blocks.%d." % l
        h = F.layer_norm(x, (C,), st[b + "norm1.weight"], st[b + "norm1.bias"], 1e-6)
        qkv = F.linear(h, st[b + "attn.qkv.weight"], st[b + "attn.qkv.bias"])
        q, k, v = qkv.reshape(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
        att = ((q @ k.transpose(-2, -1)) * D ** -0.5).softmax(-1)
        o = (att @ v).transpose(1, 2).reshape(B, N, C)
        x = x + F.linear(o, st[b + "attn.proj.weight"], st[b + "attn.proj.bias"])
        h = F.layer_norm(x, (C,), st[b + "norm2.weight"], st[b + "norm2.bias"], 1e-6)
        u = F.gelu(F.linear(h, st[b + "mlp.fc1.weight"], st[b + "mlp.fc1.bias"]))
        x = x + F.linear(u, st[b + "mlp.fc2.weight"], st[b + "mlp.fc2.bias"])
    x = F.layer_norm(x, (C,), st["norm.weight"], st["norm.bias"], 1e-6)
    return F.linear(x[:, 0], st["head.weight"], st["head.bias"])
