#!/usr/bin/env python
"""Bring-up check of the tcgen05 GEMM through the C ABI against torch (GPU) on seeded inputs."""
import ctypes as C
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import _lib as L  # noqa: E402


def run(M, N, K0, K1=0, slices=1, bias=False, epi=L.EPI_NONE, seed=0, time_it=False, tag=""):
    torch.manual_seed(seed)
    dev = "cuda"
    A0 = (torch.randn(M, K0, device=dev) * 0.5).bfloat16()
    B0 = (torch.randn(N, K0, device=dev) * 0.05).bfloat16()
    d = L.GemmDesc()
    d.M, d.N, d.K0 = M, N, K0
    d.A0, d.lda0, d.B0, d.ldb0 = A0.data_ptr(), K0, B0.data_ptr(), K0
    ref = A0.float() @ B0.float().T
    if K1:
        A1 = (torch.randn(M, slices * K1, device=dev) * 0.3).bfloat16()
        B1 = (torch.randn(N // slices, K1, device=dev) * 0.3).bfloat16()
        d.K1, d.ext_slices = K1, slices
        d.A1, d.lda1, d.B1, d.ldb1 = A1.data_ptr(), slices * K1, B1.data_ptr(), K1
        w = N // slices
        for s in range(slices):
            ref[:, s * w:(s + 1) * w] += A1[:, s * K1:(s + 1) * K1].float() @ B1.float().T
    if bias:
        b = torch.randn(N, device=dev)
        d.bias = b.data_ptr()
        ref += b
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    d.out, d.ldo = out.data_ptr(), N
    out2 = aux = None
    if epi == L.EPI_GELU:
        out2 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        d.out2, d.ldo2 = out2.data_ptr(), N
    if epi == L.EPI_DGELU:
        aux = torch.randn(M, N, device=dev).bfloat16()
        d.aux, d.ldaux = aux.data_ptr(), N
        u = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(u).sum().backward()
        ref = ref * u.grad
    d.epi = epi
    d.pair = int(os.environ.get('CARA_GEMM_PAIR', '0'))
    st = torch.cuda.current_stream().cuda_stream
    L.check(L.lib().cara_gemm_cp(C.byref(d), st), "cara_gemm_cp")
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    msg = "%-28s M=%d N=%d K0=%d K1=%d S=%d: max|err|=%.3e (max|ref|=%.2f) rel=%.3e" % (
        tag, M, N, K0, K1, slices, err, scale, rel)
    ok = rel < 6e-3 and not torch.isnan(out.float()).any().item()
    if epi == L.EPI_GELU:
        ref2 = torch.nn.functional.gelu(out.float())
        rel2 = ((out2.float() - ref2).norm() / ref2.norm()).item()
        msg += " gelu rel=%.3e" % rel2
        ok = ok and rel2 < 6e-3
    if not ok:
        bad = ((out.float() - ref).abs() > 0.05 * scale + 1e-2).nonzero()
        msg += "  FIRST BAD %s of %d; nan=%d" % (bad[:4].tolist(), bad.shape[0], int(torch.isnan(out.float()).sum()))
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    if time_it:
        for _ in range(3):
            L.lib().cara_gemm_cp(C.byref(d), st)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            L.lib().cara_gemm_cp(C.byref(d), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * M * N * (K0 + K1)
        print("     time %.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9), flush=True)
        t = torch.cuda.Event(True); t2 = torch.cuda.Event(True)
        t.record()
        for _ in range(10):
            ref = A0 @ B0.T
        t2.record(); torch.cuda.synchronize()
        print("     cuBLAS bf16 same shape: %.3f ms %.1f TFLOP/s" % (t.elapsed_time(t2) / 10, 2.0 * M * N * K0 / (t.elapsed_time(t2) / 10) / 1e9))
    return ok


if __name__ == "__main__":
    L.check(L.lib().cara_set_device(0), "set_device")
    ok = True
    ok &= run(128, 256, 64, tag="1 tile 1 kblock")
    ok &= run(128, 256, 256, tag="1 tile 4 kblocks")
    ok &= run(128, 256, 768, tag="1 tile 12 kblocks (ring wrap)")
    ok &= run(512, 1024, 768, tag="16 tiles")
    ok &= run(591, 768, 768, bias=True, tag="M tail + bias")
    ok &= run(50432, 768, 768, bias=True, tag="many tiles/CTA")
    ok &= run(1024, 2304, 768, K1=16, slices=3, bias=True, tag="qkv-like ext r16")
    ok &= run(1024, 3072, 768, K1=32, slices=4, bias=True, tag="fc1-like ext r32")
    ok &= run(1024, 768, 3072, K1=16, slices=1, bias=True, tag="fc2-like ext")
    ok &= run(640, 3072, 768, K1=16, slices=4, bias=True, epi=L.EPI_GELU, tag="fc1 GELU epilogue")
    ok &= run(640, 3072, 768, K1=16, slices=1, epi=L.EPI_DGELU, tag="dGELU epilogue")
    ok &= run(50432, 3072, 768, K1=16, slices=4, bias=True, time_it=True, tag="c2 fc1 full size")
    ok &= run(50432, 768, 3072, K1=16, slices=1, bias=True, time_it=True, tag="c2 fc2 full size")
    ok &= run(50432, 2304, 768, K1=16, slices=3, bias=True, time_it=True, tag="c2 qkv full size")
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)
