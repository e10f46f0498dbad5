"""CaRA adapter module -- drop-in for the reference's ``src/cara/cara.py`` on B200.

Same public surface as the reference (``from src.cara.cara import cara``; ``cara(config)``, ``set_cara``,
``cp_attn``, ``cp_mlp``, the module global ``global_model``, the twelve ``CP_*`` parameters on the root ViT,
``.dp .s .dim .idx .attn_idx`` on the patched children -- reference cara.py:12-188), but the two patched
forwards run on the cara_b200 CUDA kernels: the frozen projection and the rank-R chain
``x A diag(c) B^T`` are accumulated in one tcgen05 GEMM and the delta weights the reference materialises
with ``tl.cp_to_tensor`` (cara.py:27,52,76,88) are never formed.

One documented deviation: the reference applies ``nn.Dropout(0.1)`` to the *materialised delta weights*
in train mode (cara.py:35,57,81,92).  That term is full-rank and cannot be evaluated from the factors, so
``self.dp`` is kept as an attribute but not applied (hard part 1).
"""
import warnings
from typing import Any, Dict

import torch as th
import torch.nn as nn

from cara_b200 import staging as _staging
from cara_b200 import vit as _vit
from cara_b200 import wdrop as _wdrop

global_model: th.nn.Module

_ROOT_TYPES = [_vit.VisionTransformer]
_ATTN_TYPES = [_vit.Attention]
_MLP_TYPES = [_vit.Mlp]
_warned = [False]


def _weight_dropout_note(mod):
    if mod.training and mod.dp.p > 0.0 and not _warned[0]:
        _warned[0] = True
        warnings.warn("cara_b200: weight-space dropout on the CP delta (reference cara.py:35,57,81,92) is not "
                      "applied by the fused kernels; training proceeds without it "
                      "(cara_b200.wdrop.set_weight_dropout(model, 'exact') selects the exact slow path)")


def cp_attn(self, x: th.Tensor) -> th.Tensor:
    """Attention with CP parameters (reference cara.py:15-60).

    Args:
        x (th.Tensor): Input tensor [B, N, C] on a CUDA device.

    Returns:
        th.Tensor: CaRA attention output [B, N, C].
    """
    amap, _ = _staging.staged(global_model)
    if _wdrop.wants_exact(global_model, self):     # opt-in slow path: the reference's dropout on the delta weights
        return _wdrop.attn_forward(self, x, amap[id(self)])
    _weight_dropout_note(self)
    return _vit.attn_forward(self, x, amap[id(self)])


def cp_mlp(self, x: th.Tensor) -> th.Tensor:
    """Mlp with CP parameters (reference cara.py:63-95).

    Args:
        x (th.Tensor): Input tensor [B, N, C] on a CUDA device.

    Returns:
        th.Tensor: Mlp projected output [B, N, C].
    """
    _, mmap = _staging.staged(global_model)
    if _wdrop.wants_exact(global_model, self):
        return _wdrop.mlp_forward(self, x, mmap[id(self)])
    _weight_dropout_note(self)
    return _vit.mlp_forward(self, x, mmap[id(self)])


def _geometry(model: nn.Module):
    """(3L, C, H, D, 9L) generalising the reference's hard-coded 36/768/12/64/108 (cara.py:112-125)."""
    C = model.embed_dim
    depth = len(model.blocks)
    heads = model.blocks[0].attn.num_heads
    if model.blocks[0].mlp.fc1.out_features != 4 * C:
        raise ValueError("CaRA's FFN tensor needs mlp hidden == 4 * embed_dim")
    return 3 * depth, C, heads, C // heads, 9 * depth


def set_cara(model: nn.Module, rank: int, scale: float, l_mu: float, l_std: float) -> None:
    """Cara setup (reference cara.py:98-166).

    Args:
        model (nn.Module): ViT model.
        rank (int): FT Rank.
        scale (float): FT scale.
        l_mu (float): Init lambda_mu.
        l_std (float): Init lambda_std.
    """
    if type(model) in _ROOT_TYPES:
        a1_rows, C, H, D, p1_rows = _geometry(model)
        dev = model.cls_token.device
        shapes = {"CP_A1": [a1_rows, rank], "CP_A2": [C, rank], "CP_A3": [H, rank], "CP_A4": [D, rank],
                  "CP_P1": [p1_rows, rank], "CP_P2": [C, rank], "CP_P3": [C, rank], "CP_R1": [rank],
                  "CP_R2": [rank], "CP_bias1": [C], "CP_bias2": [4 * C], "CP_bias3": [C]}
        for name, shape in shapes.items():
            setattr(model, name, nn.Parameter(th.empty(shape, device=dev), requires_grad=True))
        nn.init.xavier_normal_(model.CP_A1)
        nn.init.zeros_(model.CP_A2)
        nn.init.orthogonal_(model.CP_A3)
        nn.init.orthogonal_(model.CP_A4)
        nn.init.xavier_normal_(model.CP_P1)
        nn.init.zeros_(model.CP_P2)
        nn.init.orthogonal_(model.CP_P3)
        if l_std != 0.0:
            nn.init.normal_(model.CP_R1, mean=l_mu, std=l_std)
            nn.init.normal_(model.CP_R2, mean=l_mu, std=l_std)
        elif l_mu == 1.0 and l_std == 0.0:
            nn.init.ones_(model.CP_R1)
            nn.init.ones_(model.CP_R2)
        else:
            # the reference leaves CP_R1/CP_R2 uninitialised here (cara.py:134-139); use l_mu instead of garbage
            nn.init.constant_(model.CP_R1, l_mu)
            nn.init.constant_(model.CP_R2, l_mu)
        nn.init.zeros_(model.CP_bias1)
        nn.init.zeros_(model.CP_bias2)
        nn.init.zeros_(model.CP_bias3)
        model.idx = 0
        model.attn_idx = 0
    for child in model.children():
        if type(child) in _ATTN_TYPES:
            child.dp = nn.Dropout(0.1)
            child.s = scale
            child.dim = rank
            child.idx = global_model.idx
            child.attn_idx = global_model.attn_idx
            global_model.idx += 1
            global_model.attn_idx += 3
            setattr(child, "forward", cp_attn.__get__(child, child.__class__))  # noqa: B010
        elif type(child) in _MLP_TYPES:
            child.dp = nn.Dropout(0.1)
            child.s = scale
            child.dim = rank
            child.idx = global_model.idx
            global_model.idx += 8
            setattr(child, "forward", cp_mlp.__get__(child, child.__class__))  # noqa: B010
        elif len(list(child.children())) != 0:
            set_cara(child, rank, scale, l_mu, l_std)


def cara(config: Dict[str, Any]) -> th.nn.Module:
    """Set CaRA for the given configuration (reference cara.py:169-188).

    Args:
        config (Dict[str, Any]): Dictionary containing CaRA configuration
            (keys ``model``, ``rank``, ``scale``, ``l_mu``, ``l_std``).

    Returns:
        th.nn.Module: CaRA model (the same instance that was passed in).
    """
    model = config["model"]
    rank = config["rank"]
    scale = config["scale"]
    l_mu = config["l_mu"]
    l_std = config["l_std"]

    global global_model
    global_model = model
    set_cara(model, rank, scale, l_mu, l_std)
    return global_model
