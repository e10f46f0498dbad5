"""Torch formulations used by the tests as comparators (never imported by the product)."""
import torch

BF16 = torch.bfloat16


def factor_operands(F, Rp):
    """fp32 factor [..., rows, R] -> (ext bf16 [..., rows, 3Rp] = [hi|hi|lo], t2 bf16 [..., 2Rp, rows] = [hi^T; lo^T]):
    what ``cara_factor_operands`` computes, written with torch ops (CPU or CUDA)."""
    F = F.detach().float().contiguous()
    Fp = torch.nn.functional.pad(F, (0, Rp - F.shape[-1]))
    hi = Fp.to(BF16)
    lo = (Fp - hi.float()).to(BF16)
    ext = torch.cat([hi, hi, lo], dim=-1).contiguous()
    t2 = torch.cat([hi.transpose(-1, -2), lo.transpose(-1, -2)], dim=-2).contiguous()
    return ext, t2


def stage_terms(P, ai, pi, mi, s_a, s_m, fb_proj, fb_fc1, fb_fc2):
    """The staging of the twelve CP_* parameters into the per-layer terms of SURVEY A.1, written with differentiable
    torch ops (what ``cara_stage_terms`` computes; autograd through it is the comparator for the kernel's backward).
    ``P``: dict of the CP_* tensors; ai / pi / mi: long [L] row indices (attn_idx, attn idx, mlp idx); s_a / s_m: [L]."""
    L, C, R = ai.shape[0], P["CP_A2"].shape[0], P["CP_A1"].shape[1]
    r3, r4 = torch.arange(3, device=ai.device), torch.arange(4, device=ai.device)
    sa, sm = s_a.view(L, 1, 1), s_m.view(L, 1, 1)
    kr = (P["CP_A3"][:, None, :] * P["CP_A4"][None, :, :]).reshape(C, R)
    cs_qkv = sa * (P["CP_R1"] * P["CP_A1"][ai[:, None] + r3])
    cs_proj = sa * (P["CP_R2"] * P["CP_P1"][pi][:, None, :])
    cs_fc1 = sm * (P["CP_R2"] * P["CP_P1"][mi[:, None] + r4])
    a_fc2 = (P["CP_P1"][mi[:, None] + 4 + r4][:, :, None, :] * P["CP_P2"][None, None]).reshape(L, 4 * C, R)
    cs_fc2 = sm * P["CP_R2"].view(1, 1, R).expand(L, 1, R)
    b_proj = fb_proj + s_a.view(L, 1) * P["CP_bias1"]
    b_fc1 = fb_fc1 + s_m.view(L, 1) * P["CP_bias2"]
    b_fc2 = fb_fc2 + s_m.view(L, 1) * P["CP_bias3"]
    return kr, cs_qkv, cs_proj, cs_fc1, a_fc2, cs_fc2, b_proj, b_fc1, b_fc2
