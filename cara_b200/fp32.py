"""fp32 mode of the CaRA hot path (north_star: "fp32 mode: logits rel-err <= 1e-4").

Same module tree, same staging of the CP factors, same math as the bf16 path -- but every
activation stays fp32 and every contraction runs on the SIMT fp32 kernels (``cara_sgemm``, ``cara_attn_f32``,
``cara_gelu_f32``, the fp32 variants of ``cara_ln_fwd/bwd``).  No tensor cores, no bf16.  It exists for parity
checks against the reference's fp32 PyTorch path, not for speed.  Select it with ``set_precision(model, "fp32")``.
"""
import torch

from . import kernels as K

F32 = torch.float32


def set_precision(model, precision):
    """'bf16' (default: tcgen05 kernels) or 'fp32' (SIMT fp32 parity mode) for every module of ``model``."""
    if precision not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    for m in model.modules():
        m.__dict__["cara_precision"] = precision
    return model


def is_fp32(mod):
    return mod.__dict__.get("cara_precision", "bf16") == "fp32"


def _colsum(G):
    ones = torch.ones((1, G.shape[0]), device=G.device, dtype=F32)
    return K.sgemm(ones, G).view(-1)


class CPLinearF32(torch.autograd.Function):
    """y = x W^T + b + sum_s (x A) Q_s^T on output slice s, Q_s = B (.) cs_s  (one CP-adapted frozen projection)."""

    @staticmethod
    def forward(ctx, x, W, bias, A, Q):
        y = K.sgemm(x, W.t(), bias=bias)
        T = None
        if A is not None:
            T = K.sgemm(x, A)
            w = Q.shape[1]
            for s in range(Q.shape[0]):
                K.sgemm(T, Q[s].t(), out=y[:, s * w:(s + 1) * w], beta=1.0)
        ctx.save_for_backward(x, W, A, Q, T)
        return y

    @staticmethod
    def backward(ctx, G):
        x, W, A, Q, T = ctx.saved_tensors
        G = G.contiguous()
        ni = ctx.needs_input_grad
        dx = K.sgemm(G, W) if ni[0] else None
        dA = dQ = None
        if A is not None:
            S, w, R = Q.shape
            dT = torch.empty((G.shape[0], R), device=G.device, dtype=F32)
            dQ = torch.empty_like(Q)
            for s in range(S):
                Gs = G[:, s * w:(s + 1) * w]
                K.sgemm(Gs, Q[s], out=dT, beta=1.0 if s else 0.0)
                K.sgemm(Gs.t(), T, out=dQ[s])
            if dx is not None:
                K.sgemm(dT, A.t(), out=dx, beta=1.0)
            dA = K.sgemm(x.t(), dT)
        dbias = _colsum(G) if (ctx.needs_input_grad[2]) else None
        return dx, None, dbias, dA, dQ


class GeluF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u):
        ctx.save_for_backward(u)
        return K.gelu_f32(u)

    @staticmethod
    def backward(ctx, dy):
        (u,) = ctx.saved_tensors
        return K.gelu_f32(u, dy=dy.contiguous())


class AttnCoreF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, B, N, H, D, scale):
        o, lse = K.attn_f32_fwd(qkv, B, N, H, D, scale, train=ctx.needs_input_grad[0])
        ctx.dims = (B, N, H, D, scale)
        ctx.save_for_backward(qkv, o, lse)
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse = ctx.saved_tensors
        B, N, H, D, scale = ctx.dims
        return K.attn_f32_bwd(qkv, o, lse, d_o.contiguous(), B, N, H, D, scale), None, None, None, None, None


def _terms(t):
    """staging.Terms -> (bias, A, Q) with Q[s] = B (.) cs_s (differentiable fp32 torch ops on factor-sized tensors)."""
    if t is None:
        return None, None, None
    return t.bias, t.A.contiguous(), (t.B[None, :, :] * t.cs[:, None, :]).contiguous()


def _lin(lin):
    return lin.weight.detach().float(), (None if lin.bias is None else lin.bias.detach().float())


def cp_linear(x, lin, t):
    W, b = _lin(lin)
    bias, A, Q = _terms(t)
    return CPLinearF32.apply(x, W, bias if bias is not None else b, A, Q)


def attn_forward(mod, x, staged):
    B, N, C = x.shape
    H = mod.num_heads
    h = x.reshape(B * N, C).float().contiguous()
    qkv = cp_linear(h, mod.qkv, None if staged is None else staged[0])
    o = AttnCoreF32.apply(qkv, B, N, H, C // H, float(mod.scale))
    y = cp_linear(o, mod.proj, None if staged is None else staged[1])
    return y.view(B, N, C)


def mlp_forward(mod, x, staged):
    B, N, C = x.shape
    h = x.reshape(B * N, C).float().contiguous()
    u = cp_linear(h, mod.fc1, None if staged is None else staged[0])
    g = GeluF32.apply(u)
    y = cp_linear(g, mod.fc2, None if staged is None else staged[1])
    return y.view(B, N, -1)


def patch_embed(mod, img):
    """fp32 im2col (a strided view: data movement only) + cara_sgemm against the flattened conv weight."""
    B = img.shape[0]
    P = mod.patch_size[0]
    w = mod.proj.weight.detach().float()
    cols = img.float().unfold(2, P, P).unfold(3, P, P)                   # [B, Cin, gh, gw, P, P]
    cols = cols.permute(0, 2, 3, 1, 4, 5).reshape(B * mod.num_patches, -1).contiguous()
    y = K.sgemm(cols, w.reshape(w.shape[0], -1).t(), bias=mod.proj.bias.detach().float())
    return y.view(B, mod.num_patches, -1)
