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
