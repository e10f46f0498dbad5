"""Per-step staging of the CP factors into the per-projection terms the kernels consume.

SURVEY Appendix A.1 table: every adapted projection is  Y = X W^T + b + s ((X A (.) c) B^T + beta).  Here the
twelve ``CP_*`` parameters of the root model (cara.py:112-125) are turned -- with differentiable fp32 torch
ops over tensors of a few thousand elements, batched over all layers -- into, per layer and projection,
``A`` [K,R], ``cs = s*c`` [slices,R], ``B`` [N/slices,R] and ``bias = b + s*beta``; autograd carries the
kernels' gradients w.r.t. these terms back to the ``CP_*`` parameters (SURVEY A.2 chain rule).  The row maps
follow the reference's own indices: ``CP_A1[attn_idx : attn_idx+3]`` (cara.py:26), ``CP_P1[idx]`` (:51),
``CP_P1[idx : idx+4]`` / ``[idx+4 : idx+8]`` (:72-73).  Results are cached until a parameter changes.
"""
import torch

from . import kernels as K
from .ops import AdapterOperands, GradSink

BF16, F32 = torch.bfloat16, torch.float32
CP_NAMES = ("CP_A1", "CP_A2", "CP_A3", "CP_A4", "CP_P1", "CP_P2", "CP_P3", "CP_R1", "CP_R2",
            "CP_bias1", "CP_bias2", "CP_bias3")


class Terms:
    """One projection's terms.  ``A cs B bias`` are this layer's tensors (fp32 parity mode, un-sunk path);
    ``stacked`` = (A or A[L,..], cs [L,S,R], B, bias [L,N] or None) are the all-layer tensors the bf16 path hands to
    autograd together with ``sink`` = (ops.GradSink, kind, layer) when gradients are on."""
    __slots__ = ("A", "cs", "B", "bias", "ops", "stacked", "sink")

    def __init__(self, A, cs, B, bias, ops, stacked=None, sink=None):
        self.A, self.cs, self.B, self.bias, self.ops = A, cs, B, bias, ops
        self.stacked, self.sink = stacked, sink

    def autograd_args(self):
        """(A, cs, B, bias, sink) for ops.CPLinearFunction / CPMlpFunction."""
        if self.sink is not None:
            return self.stacked + (self.sink,)
        return self.A, self.cs, self.B, self.bias, None


def _modules(model):
    from .vit import Attention, Mlp
    attn = [m for m in model.modules() if isinstance(m, Attention) and hasattr(m, "attn_idx")]
    mlp = [m for m in model.modules() if isinstance(m, Mlp) and hasattr(m, "idx")]
    return attn, mlp


def staged(model):
    """-> (dict id(attention module) -> (qkv Terms, proj Terms), dict id(mlp module) -> (fc1 Terms, fc2 Terms))."""
    P = {n: getattr(model, n) for n in CP_NAMES}
    attn, mlp = _modules(model)
    frozen = [m.proj.bias for m in attn] + [m.fc1.bias for m in mlp] + [m.fc2.bias for m in mlp]
    grad_on = torch.is_grad_enabled()
    token = model.__dict__.get("_cara_fwd_token")
    key = (tuple((p.data_ptr(), p._version) for p in P.values()), K.param_generation, grad_on,
           tuple(float(m.s) for m in attn + mlp), tuple((b.data_ptr(), b._version) for b in frozen),
           tuple(int(m.attn_idx) for m in attn), tuple(int(m.idx) for m in attn + mlp))
    cache = model.__dict__.get("_cara_stage")
    # With autograd on, the staged tensors carry a graph that one backward consumes: share them only among the
    # layers of ONE root forward (token set by VisionTransformer.forward_features); otherwise rebuild.
    if cache is not None and cache[0] == key and (not grad_on or (token is not None and cache[3] is token)):
        return cache[1], cache[2]
    dev = P["CP_A1"].device
    R = P["CP_A1"].shape[1]
    Rp = K.round_rank(R)
    if R > 32:
        raise NotImplementedError("rank > 32 is not supported by the staged bf16 operands (got %d)" % R)
    C = P["CP_A2"].shape[0]
    f = {n: p.float() for n, p in P.items()}
    La, Lm = len(attn), len(mlp)
    # index / scale tensors: built once per (indices, scales) -- a host list -> device copy is a synchronous
    # transfer, which must not happen inside a CUDA-graph capture of the step
    ikey = (key[3], key[5], key[6], str(dev))
    icache = model.__dict__.get("_cara_stage_idx")
    if icache is None or icache[0] != ikey:
        icache = (ikey,
                  torch.tensor([int(m.attn_idx) for m in attn], device=dev),
                  torch.tensor([int(m.idx) for m in attn], device=dev),
                  torch.tensor([int(m.idx) for m in mlp], device=dev),
                  torch.tensor([float(m.s) for m in attn], device=dev, dtype=F32).view(La, 1, 1),
                  torch.tensor([float(m.s) for m in mlp], device=dev, dtype=F32).view(Lm, 1, 1),
                  torch.arange(3, device=dev), torch.arange(4, device=dev))
        model.__dict__["_cara_stage_idx"] = icache
    _, ai, pi, mi, s_a, s_m, r3, r4 = icache

    kr_attn = (f["CP_A3"][:, None, :] * f["CP_A4"][None, :, :]).reshape(C, R)            # B of qkv
    cs_qkv = s_a * (f["CP_R1"] * f["CP_A1"][ai[:, None] + r3])                            # [La,3,R]
    cs_proj = s_a * (f["CP_R2"] * f["CP_P1"][pi][:, None, :])                             # [La,1,R]
    cs_fc1 = s_m * (f["CP_R2"] * f["CP_P1"][mi[:, None] + r4])                            # [Lm,4,R]
    a_fc2 = (f["CP_P1"][mi[:, None] + 4 + r4][:, :, None, :] * f["CP_P2"][None, None]).reshape(Lm, 4 * C, R)
    cs_fc2 = s_m * f["CP_R2"].view(1, 1, R).expand(Lm, 1, R)
    # the frozen biases never change during fine-tuning: stack them once per (pointer, version) set
    bkey = key[4]
    bcache = model.__dict__.get("_cara_stage_bias")
    if bcache is None or bcache[0] != bkey:
        bcache = (bkey, torch.stack([m.proj.bias.detach().float() for m in attn]),
                  torch.stack([m.fc1.bias.detach().float() for m in mlp]),
                  torch.stack([m.fc2.bias.detach().float() for m in mlp]))
        model.__dict__["_cara_stage_bias"] = bcache
    b_proj = bcache[1] + s_a.view(La, 1) * f["CP_bias1"]
    b_fc1 = bcache[2] + s_m.view(Lm, 1) * f["CP_bias2"]
    b_fc2 = bcache[3] + s_m.view(Lm, 1) * f["CP_bias3"]

    with torch.no_grad():
        a2_pad, a2_t = K.factor_operands(f["CP_A2"], Rp)
        kr_pad, kr_t = K.factor_operands(kr_attn, Rp)
        p2_pad, p2_t = K.factor_operands(f["CP_P2"], Rp)
        p3_pad, p3_t = K.factor_operands(f["CP_P3"], Rp)
        afc2_pad, afc2_t = K.factor_operands(a_fc2, Rp)
        pad = lambda t: torch.nn.functional.pad(t.detach(), (0, Rp - R)).contiguous()  # noqa: E731
        csq, csp, cs1, cs2 = pad(cs_qkv), pad(cs_proj), pad(cs_fc1), pad(cs_fc2)

    # gradients on: every projection's backward accumulates into one zero-filled sink (ops.GradSink)
    sink = GradSink(max(La, Lm), C, Rp, dev) if (grad_on and La == Lm) else None
    sk = (lambda kind, i: (sink, kind, i)) if sink is not None else (lambda kind, i: None)
    amap, mmap = {}, {}
    for i, m in enumerate(attn):
        qkv = Terms(f["CP_A2"], cs_qkv[i], kr_attn, None, AdapterOperands(a2_pad, a2_t, kr_pad, kr_t, csq[i], R),
                    (f["CP_A2"], cs_qkv, kr_attn, None), sk(0, i))
        proj = Terms(f["CP_P3"], cs_proj[i], f["CP_P2"], b_proj[i],
                     AdapterOperands(p3_pad, p3_t, p2_pad, p2_t, csp[i], R),
                     (f["CP_P3"], cs_proj, f["CP_P2"], b_proj), sk(1, i))
        amap[id(m)] = (qkv, proj)
    for i, m in enumerate(mlp):
        fc1 = Terms(f["CP_P3"], cs_fc1[i], f["CP_P2"], b_fc1[i], AdapterOperands(p3_pad, p3_t, p2_pad, p2_t, cs1[i], R),
                    (f["CP_P3"], cs_fc1, f["CP_P2"], b_fc1), sk(2, i))
        fc2 = Terms(a_fc2[i], cs_fc2[i], f["CP_P3"], b_fc2[i],
                    AdapterOperands(afc2_pad[i], afc2_t[i], p3_pad, p3_t, cs2[i], R),
                    (a_fc2, cs_fc2, f["CP_P3"], b_fc2), sk(3, i))
        mmap[id(m)] = (fc1, fc2)
    model.__dict__["_cara_stage"] = (key, amap, mmap, token)
    return amap, mmap
