#!/usr/bin/env python
"""Isolated time of the fused projection GEMM with and without its side tiles at the bench shapes
(ViT-B/16, batch 256: M = 50,432, rank 16), CUDA events, L2 flushed between launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import _lib as L, kernels as K

M = int(os.environ.get("M", 50432))
R = int(os.environ.get("R", 16))
Rp = K.round_rank(R)
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def t(fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


# (name, N, K0, S_ext, epi, side mode, side slices)
SHAPES = [("qkv fwd", 2304, 768, 3, 0, L.SIDE_FWD, 3), ("proj fwd", 768, 768, 1, 0, L.SIDE_FWD, 1),
          ("fc1 fwd gelu", 3072, 768, 4, 1, L.SIDE_FWD, 4), ("fc2 fwd", 768, 3072, 1, 0, L.SIDE_FWD, 1),
          ("fc2 dx dgelu", 3072, 768, 1, 2, L.SIDE_BWD, 1), ("fc1 dx", 768, 3072, 1, 0, L.SIDE_BWD, 4),
          ("proj dx", 768, 768, 1, 0, L.SIDE_BWD, 1), ("qkv dx", 768, 2304, 1, 0, L.SIDE_BWD, 3)]
for name, N, K0, S, epi, mode, ss in SHAPES:
    x = (torch.randn(M, K0, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K0, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    kslices = ss if mode == L.SIDE_BWD else 1
    P = (torch.randn(2 * Rp, K0 // kslices, device=dev, generator=g) * 0.1).to(torch.bfloat16)
    cs = torch.randn(ss, Rp, device=dev, generator=g)
    T = torch.randn(M, Rp, device=dev, generator=g)
    U = (torch.randn(M, (3 * Rp) if mode == L.SIDE_BWD else ss * 3 * Rp, device=dev, generator=g) * 0.1).to(torch.bfloat16)
    b1 = (torch.randn(N // S, 3 * Rp, device=dev, generator=g) * 0.1).to(torch.bfloat16)
    dc = torch.zeros(ss, Rp, device=dev)
    aux = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16) if epi == 2 else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi == 1 else None
    side = K.Side(mode, P, cs, T, U, dc if mode == L.SIDE_BWD else None)
    plain = t(lambda: K.gemm_cp(x, W, a1=U, b1=b1, ext_slices=S, epi=epi, aux=aux, out=out, out2=out2))
    if epi != 0:                      # the GELU kinds have no side-drain warps: stand-alone rows pass only
        fused = float("nan")
    else:
        fused = t(lambda: K.gemm_cp(x, W, a1=U, b1=b1, ext_slices=S, epi=epi, aux=aux, out=out, out2=out2, side=side))
    if mode == L.SIDE_FWD:
        alone = t(lambda: K.adapter_rows_fwd(x, P, cs))
    else:
        alone = t(lambda: K.adapter_rows_bwd(x, P, cs, T))
    fl = 2.0 * M * N * (K0 + Rp)
    print("%-14s N%5d K%5d  plain %7.1f us (%6.1f TF)  +side %7.1f us (+%5.1f)  side tiles alone %6.1f us"
          % (name, N, K0, plain, fl / plain / 1e6, fused, fused - plain, alone), flush=True)
    del x, W, out, out2, aux
    torch.cuda.empty_cache()
