"""Exact weight-space dropout -- the reference's train-mode semantics, as an opt-in slow path.

In train mode the reference applies ``self.dp = nn.Dropout(0.1)`` to the MATERIALISED delta weights before using
them (cara.py:35,57,81,92):  y = x W^T + b + s (x (M (.) dW)^T + beta),  M Bernoulli(0.9) / 0.9 per weight element.
``M (.) dW`` is full rank, so it cannot ride in the rank-R adapter segment of the fused projection, and its gradient
needs the dense dW-shaped product G^T X that the fast path is built to avoid.  The fast path therefore runs without
it (documented deviation, one warning).  ``set_weight_dropout(model, "exact")`` selects this module instead:

  forward   dW = (B (.) cs) A^T  [N,K] fp32 (cara_sgemm), mask, W_eff = W + M (.) dW  (bf16, both layouts),
            y = x W_eff^T + b_eff                       one tcgen05 GEMM (fc1: GELU epilogue), no adapter segment
  backward  dX = G W_eff                                one tcgen05 GEMM (fc2: GELU' epilogue)
            dDelta = M (.) (G^T X)                      one more tcgen05 GEMM over transposed copies of G and X
            dA = dDelta^T (B (.) cs),  d(B (.) cs) = dDelta A    (cara_sgemm), then the usual chain to the CP factors

i.e. three big GEMMs per projection instead of two plus skinny contractions (about 1.5x the step time).  With p = 0 or
in eval mode the fast path is used regardless.  Masks are drawn with torch's CUDA generator (the reference's
CPU / CUDA generator streams cannot be reproduced bit for bit anyway); ``MASK_TAP`` lets the tests replay them
through the oracle.
"""
import torch

from . import _lib as L
from . import kernels as K
from . import ops

BF16, F32 = torch.bfloat16, torch.float32
MASK_TAP = None            # tests: a list that receives every drawn mask ([N,K], already scaled by 1/(1-p))


def set_weight_dropout(model, mode):
    """'skip' (default: fused fast path, dropout on the delta not applied) or 'exact' (this module)."""
    if mode not in ("skip", "exact"):
        raise ValueError("mode must be 'skip' or 'exact'")
    model.__dict__["cara_weight_dropout"] = mode
    return model


def wants_exact(root, mod):
    return mod.training and mod.dp.p > 0.0 and root.__dict__.get("cara_weight_dropout", "skip") == "exact"


def _pad_rows(t, mult):
    """bf16 [R, M] -> M padded with zeros to a multiple of ``mult`` (the GEMM's K granularity)."""
    M = t.shape[1]
    Mp = (M + mult - 1) // mult * mult
    return t if Mp == M else torch.nn.functional.pad(t, (0, Mp - M))


def _effective(lin, t, p):
    """-> (W_eff bf16 [N,K], W_eff^T bf16 [K,N], mask fp32 [N,K], Q fp32 [N,R] = B (.) cs stacked over slices)."""
    W = lin.weight.detach().float()
    N, Kin = W.shape
    A, Bf, cs = t.A.detach().float().contiguous(), t.B.detach().float(), t.cs.detach().float()
    Q = (Bf[None, :, :] * cs[:, None, :]).reshape(N, -1).contiguous()
    dW = K.sgemm(Q, A.t())                                         # [N,K] fp32: the materialised delta (times s)
    mask = (torch.rand((N, Kin), device=W.device, dtype=F32) >= p).to(F32) / (1.0 - p)
    if MASK_TAP is not None:
        MASK_TAP.append(mask.detach().clone())
    w_eff = torch.addcmul(W, dW, mask)
    return w_eff.to(BF16).contiguous(), w_eff.t().to(BF16).contiguous(), mask, Q


def _factor_grads(G, x, mask, A, Q, t, need_bias):
    """dDelta = M (.) (G^T X) -> (dA [K,R], dcs [S,R], dB [w,R], dbias [N]) for the staged terms."""
    Gt = _pad_rows(G.t().contiguous(), 64)
    Xt = _pad_rows(x.t().contiguous(), 64)
    d_delta = K.gemm_cp(Gt, Xt).float() * mask                      # [N,K]
    dA = K.sgemm(d_delta.t(), Q)                                    # [K,R]
    dQ = K.sgemm(d_delta, A)                                        # [N,R]
    S = t.cs.shape[0]
    dQs = dQ.view(S, -1, dQ.shape[1])
    dB = (dQs * t.cs.detach().float()[:, None, :]).sum(0)
    dcs = (dQs * t.B.detach().float()[None, :, :]).sum(1)
    dbias = G.float().sum(0) if need_bias else None
    return dA, dcs, dB, dbias


class _Saved:
    __slots__ = ("wt", "mask", "A", "Q", "t")


def _fwd(x, lin, t, p, epi=L.EPI_NONE, want_pre=True):
    w, wt, mask, Q = _effective(lin, t, p)
    bias = t.bias.detach().float().contiguous() if t.bias is not None else ops.FrozenLinear.of(lin).bias
    y = K.gemm_cp(x, w, bias=bias, epi=epi, want_pre=want_pre)
    s = _Saved()
    s.wt, s.mask, s.A, s.Q, s.t = wt, mask, t.A.detach().float().contiguous(), Q, t
    return y, s


class DropLinearFunction(torch.autograd.Function):
    """One projection with exact weight dropout on its CP delta (qkv: cara.py:25-42, proj: :50-58)."""

    @staticmethod
    def forward(ctx, x, A, cs, Bf, bias_eff, lin, t, p):
        y, ctx.s = _fwd(x, lin, t, p)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, G):
        (x,) = ctx.saved_tensors
        G = G.contiguous()
        s = ctx.s
        dx = K.gemm_cp(G, s.wt) if ctx.needs_input_grad[0] else None
        dA, dcs, dB, dbias = _factor_grads(G, x, s.mask, s.A, s.Q, s.t, ctx.needs_input_grad[4])
        return dx, dA, dcs, dB, dbias, None, None, None


class DropMlpFunction(torch.autograd.Function):
    """cp_mlp (cara.py:72-95) with exact weight dropout: GELU in the fc1 GEMM epilogue, GELU' in the fc2 dX epilogue."""

    @staticmethod
    def forward(ctx, x, A1, cs1, B1, bias1, A2, cs2, B2, bias2, mod, t1, t2, p):
        (gp, g), ctx.s1 = _fwd(x, mod.fc1, t1, p, epi=L.EPI_GELU)   # gp = gelu'(u), g = GELU(u)
        y, ctx.s2 = _fwd(g, mod.fc2, t2, p)
        ctx.save_for_backward(x, gp, g)
        return y

    @staticmethod
    def backward(ctx, G):
        x, gp, g = ctx.saved_tensors
        G = G.contiguous()
        ni = ctx.needs_input_grad
        s1, s2 = ctx.s1, ctx.s2
        du = K.gemm_cp(G, s2.wt, epi=L.EPI_DGELU, aux=gp)
        dA2, dcs2, dB2, db2 = _factor_grads(G, g, s2.mask, s2.A, s2.Q, s2.t, ni[8])
        dx = K.gemm_cp(du, s1.wt) if ni[0] else None
        dA1, dcs1, dB1, db1 = _factor_grads(du, x, s1.mask, s1.A, s1.Q, s1.t, ni[4])
        return dx, dA1, dcs1, dB1, db1, dA2, dcs2, dB2, db2, None, None, None, None


def attn_forward(mod, x, staged):
    """cp_attn with exact weight dropout; the attention core is the fast path's kernel."""
    from .vit import _as_act, _require_cuda
    _require_cuda(x, "Attention.forward")
    h, (B, N, C) = _as_act(x)
    tq, tp = staged
    p = float(mod.dp.p)
    qkv = DropLinearFunction.apply(h, tq.A, tq.cs, tq.B, None, mod.qkv, tq, p)
    o = ops.AttnCoreFunction.apply(qkv, B, N, mod.num_heads, C // mod.num_heads, float(mod.scale))
    y = DropLinearFunction.apply(o, tp.A, tp.cs, tp.B, tp.bias, mod.proj, tp, p).view(B, N, C)
    return y if x.dtype == BF16 else y.to(x.dtype)


def mlp_forward(mod, x, staged):
    from .vit import _as_act, _require_cuda
    _require_cuda(x, "Mlp.forward")
    h, (B, N, C) = _as_act(x)
    t1, t2 = staged
    y = DropMlpFunction.apply(h, t1.A, t1.cs, t1.B, t1.bias, t2.A, t2.cs, t2.B, t2.bias, mod, t1, t2,
                              float(mod.dp.p)).view(B, N, -1)
    return y if x.dtype == BF16 else y.to(x.dtype)
