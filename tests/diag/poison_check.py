#!/usr/bin/env python
"""(test infrastructure, not collected by pytest)  Does any kernel read memory it did not write?  Each op runs once on clean memory and once after the caching
allocator's free blocks were filled with NaN; outputs must be identical (up to the fp32 atomics' order)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cara_b200 import kernels as K, _lib as L
BF16, F32 = torch.bfloat16, torch.float32

def poison():
    torch.cuda.empty_cache()
    big = [torch.full((256 << 20,), float("nan"), device="cuda") for _ in range(4)]
    small = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 8, 1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22) for _ in range(24)]
    del big, small
    torch.cuda.synchronize()

def flat(o):
    if o is None: return []
    if isinstance(o, torch.Tensor): return [o]
    return [t for x in o for t in flat(x)]

def check(name, fn):
    torch.cuda.synchronize()
    a = [t.clone() for t in flat(fn())]
    poison()
    b = flat(fn())
    torch.cuda.synchronize()
    worst = 0.0
    for x, y in zip(a, b):
        if not torch.isfinite(y.float()).all():
            worst = float("inf"); break
        d = (x.float() - y.float()).abs().max().item() / max(x.float().abs().max().item(), 1e-30)
        worst = max(worst, d)
    print("%-28s max rel diff clean vs poisoned: %.2e %s" % (name, worst, "<-- reads uninitialised memory" if worst > 1e-5 else ""))

g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
B, N, H, D, C, R, Rp = 4, 197, 12, 64, 768, 8, 16
M = B * N
x = rn(M, C).to(BF16); w = (rn(3 * C, C) * 0.03).to(BF16); bias = rn(3 * C)
A = rn(C, R) * 0.1; Bf = rn(C, R) * 0.1
a_ext, a_t2 = K.factor_operands(A, Rp); b_ext, b_t2 = K.factor_operands(Bf, Rp)
cs = torch.nn.functional.pad(rn(3, R), (0, Rp - R)).contiguous()
T, U = K.adapter_rows_fwd(x, a_t2, cs)
check("rows_fwd", lambda: K.adapter_rows_fwd(x, a_t2, cs))
check("gemm + adapter segment", lambda: K.gemm_cp(x, w, bias=bias, a1=U, b1=b_ext, ext_slices=3))
w4 = (rn(4 * C, C) * 0.03).to(BF16); b4 = rn(4 * C)
check("gemm GELU epilogue", lambda: K.gemm_cp(x, w4, bias=b4, epi=L.EPI_GELU))
u = rn(M, 4 * C).to(BF16); G4 = rn(M, 4 * C).to(BF16); w4t = (rn(C, 4 * C) * 0.03).to(BF16)
Gc = rn(M, C).to(BF16); wt = (rn(4 * C, C) * 0.03).to(BF16)
check("gemm GELU' epilogue", lambda: K.gemm_cp(Gc, wt, epi=L.EPI_DGELU, aux=u))
qkv = K.gemm_cp(x, w, bias=bias)
scale = D ** -0.5
o, o_lo, lse = K.attn_fwd(qkv.view(-1), B, N, H, D, scale)
check("attn_fwd", lambda: K.attn_fwd(qkv.view(-1), B, N, H, D, scale))
d_o = rn(M, C).to(BF16)
check("attn_bwd", lambda: K.attn_bwd(qkv.view(-1), o, o_lo, lse, d_o, B, N, H, D, scale))
G3 = rn(M, 3 * C).to(BF16)
check("rows_bwd", lambda: K.adapter_rows_bwd(G3, b_t2, cs, T))
dT, _ = K.adapter_rows_bwd(G3, b_t2, cs, T)
check("cols dA", lambda: K.adapter_cols(x, dT, 1, Rp))
check("cols dB + colsum", lambda: K.adapter_cols(G3, U, 3, Rp, want_colsum=True))
xr = rn(M, C); gamma = rn(C); beta = rn(C); delta = rn(M, C).to(BF16); rs = torch.ones(B, device="cuda")
check("ln_fwd", lambda: K.ln_fwd(xr, gamma, beta, delta=delta, rowscale=rs, rows_per_sample=N))
xo, h, mean, rstd = K.ln_fwd(xr, gamma, beta, delta=delta, rowscale=rs, rows_per_sample=N)
dh = rn(M, C).to(BF16); dxin = rn(M, C)
check("ln_bwd", lambda: K.ln_bwd(dh, xo, mean, rstd, gamma, dx_in=dxin, rowscale=rs, rows_per_sample=N, want_g=True))
img = rn(B, 3, 224, 224)
check("patchify", lambda: K.patchify(img, 16, 768))
