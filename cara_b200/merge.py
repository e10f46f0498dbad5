"""Eval-mode merge: fold the CP delta into the frozen weights once, then run the plain (un-adapted) kernels.

The reference never merges -- in eval it re-materialises all four delta tensors per block on every forward
(cara.py:27,52,76,88; SURVEY call stack 2).  SURVEY A.3:  W_eff = W + s * dW,  b_eff = b + s * beta, with
dW given by the staged (A, cs, B) of each projection.  ``merge_cara`` writes the merged bf16 weights with the
``cara_merge_weights`` kernel (HBM-bound: 12 C^2 L elements read + written) and restores the un-adapted
forwards, so ``model(x)`` costs exactly the frozen ViT.
"""
import torch

from . import kernels as K
from . import ops, staging


def _merged(lin, t):
    W = lin.weight.detach()
    if W.dtype != torch.float32 or not W.is_contiguous():
        W = W.float().contiguous()
    out = ops.FrozenLinear.__new__(ops.FrozenLinear)
    out.w = K.merge_weights(W, t.A.detach(), t.B.detach(), t.cs.detach())
    out.wt = None                    # inference only: no dX operand
    if t.bias is not None:
        out.bias = t.bias.detach().float().contiguous()
    else:
        out.bias = None if lin.bias is None else lin.bias.detach().float().contiguous()
    out.key = ops.FrozenLinear.key_of(lin)
    return out


@torch.no_grad()
def merge_cara(model):
    """Fold every block's CP delta into its four projections (in the bf16 working copies; the fp32 parameters
    and the state_dict stay untouched) and switch the patched modules back to the plain forwards."""
    amap, mmap = staging.staged(model)
    from .vit import Attention, Mlp
    n = 0
    for m in model.modules():
        if isinstance(m, Attention) and id(m) in amap:
            q, p = amap[id(m)]
            m.qkv.__dict__["_cara_frozen"] = _merged(m.qkv, q)
            m.proj.__dict__["_cara_frozen"] = _merged(m.proj, p)
            m.__dict__.pop("forward", None)
            n += 2
        elif isinstance(m, Mlp) and id(m) in mmap:
            u, d = mmap[id(m)]
            m.fc1.__dict__["_cara_frozen"] = _merged(m.fc1, u)
            m.fc2.__dict__["_cara_frozen"] = _merged(m.fc2, d)
            m.__dict__.pop("forward", None)
            n += 2
    model.eval()
    return n
