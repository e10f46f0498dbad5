"""CPU tests of the host-side logic: C-ABI exports, factor staging vs. the oracle's factoring, checkpoint
schema round trip with the oracle's state, launcher argument checks."""
import os
import re

import pytest
import torch

from oracle import cara_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    import ctypes
    from cara_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "cara_b200.h")).read()
    declared = sorted(set(re.findall(r"CARA_API\s+[\w\s\*]+?\b(cara_\w+)\s*\(", hdr)))
    assert len(declared) >= 16
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(L.exported_symbols()) == declared
    assert L.lib().cara_abi_version() == L.ABI_VERSION


def _small():
    from cara_b200.vit import create_model
    from src.cara.cara import cara
    g = O.Geometry(embed_dim=256, depth=2, num_heads=4, rank=8, num_classes=7)
    vit = create_model("vit_base_patch16_224_in21k", depth=2, embed_dim=256, num_heads=4)
    vit = cara({"model": vit, "rank": 8, "scale": 2.5, "l_mu": 1.0, "l_std": 0.0})
    vit.reset_classifier(7)
    st = O.synthetic_state(g, dtype=torch.float32)
    vit.load_state_dict(st, strict=True)
    return vit, st, g


def test_torch_staging_reference_matches_oracle_factoring():
    """tests/_torch_ref.stage_terms -- the torch formulation the GPU test checks ``cara_stage_terms`` (and its backward)
    against -- reproduces the oracle's factoring of every projection on the CPU."""
    from tests import _torch_ref
    g = O.Geometry(embed_dim=256, depth=2, num_heads=4, rank=8, num_classes=7)
    st = O.synthetic_state(g, dtype=torch.float32)
    L_, C = g.depth, g.embed_dim
    s = 2.5
    P = {k: st[k] for k in st if k.startswith("CP_")}
    ai, pi, mi = torch.arange(L_) * 3, torch.arange(L_) * 9, torch.arange(L_) * 9 + 1      # cara.py:150-162 row maps
    sc = torch.full((L_,), s)
    fb = [torch.stack([st["blocks.%d.%s.bias" % (l, n)] for l in range(L_)]) for n in ("attn.proj", "mlp.fc1", "mlp.fc2")]
    kr, cs_qkv, cs_proj, cs_fc1, a_fc2, cs_fc2, b_proj, b_fc1, b_fc2 = _torch_ref.stage_terms(P, ai, pi, mi, sc, sc, *fb)
    for l in range(L_):
        for which, A_, cs_, B_, bias_ in (("qkv", st["CP_A2"], cs_qkv[l], kr, None), ("proj", st["CP_P3"], cs_proj[l], st["CP_P2"], b_proj[l]),
                                          ("fc1", st["CP_P3"], cs_fc1[l], st["CP_P2"], b_fc1[l]),
                                          ("fc2", a_fc2[l], cs_fc2[l], st["CP_P3"], b_fc2[l])):
            A, c, B, beta = O.adapter_terms(st, g, l, which)
            assert torch.allclose(A_, A, atol=1e-6) and torch.allclose(B_, B, atol=1e-6) and torch.allclose(cs_, s * c, atol=1e-6), which
            if beta is not None:
                key = {"proj": "attn.proj", "fc1": "mlp.fc1", "fc2": "mlp.fc2"}[which]
                assert torch.allclose(bias_, st["blocks.%d.%s.bias" % (l, key)] + s * beta, atol=1e-6)


@pytest.mark.gpu
def test_staging_matches_oracle_factoring_and_grad_chain():
    """staged (A, cs, B, bias) == oracle.adapter_terms (A.1) and autograd through the staging (one ``cara_stage_terms``
    launch each way) reproduces the oracle's chain rule to the CP parameters (A.2); operand layouts; cache behaviour."""
    from cara_b200 import staging
    vit, st, g = _small()
    vit = vit.cuda()
    st = {k: v.cuda() for k, v in st.items()}
    amap, mmap = staging.staged(vit)
    s = 2.5
    torch.manual_seed(0)
    total = 0.0
    ref_total = 0.0
    leaves = {k: st[k].clone().requires_grad_(True) for k in st if k.startswith("CP_")}
    work = dict(st); work.update(leaves)
    for l, blk in enumerate(vit.blocks):
        for which, t in (("qkv", amap[id(blk.attn)][0]), ("proj", amap[id(blk.attn)][1]),
                         ("fc1", mmap[id(blk.mlp)][0]), ("fc2", mmap[id(blk.mlp)][1])):
            A, c, B, beta = O.adapter_terms(work, g, l, which)
            assert torch.allclose(t.A, A.detach(), atol=1e-6) and torch.allclose(t.B, B.detach(), atol=1e-6)
            assert torch.allclose(t.cs, (s * c).detach(), atol=1e-6), which
            assert t.ops.cs_pad.shape == (c.shape[0], 16) and t.ops.a_t2.shape == (32, A.shape[0])
            assert torch.allclose(t.ops.cs_pad[:, :8], (s * c).detach(), atol=1e-6) and float(t.ops.cs_pad[:, 8:].abs().max()) == 0
            hi = A.detach().bfloat16()
            assert torch.equal(t.ops.a_ext[:, :8], hi) and torch.equal(t.ops.a_ext[:, 16:24], hi)
            assert torch.equal(t.ops.a_ext[:, 32:40], (A.detach() - hi.float()).bfloat16())
            assert float(t.ops.a_ext[:, 8:16].abs().max()) == 0 and torch.equal(t.ops.a_t2[:16].t(), t.ops.a_ext[:, :16])
            recon = t.ops.a_ext[:, :8].float() + t.ops.a_ext[:, 32:40].float()
            assert float((recon - A.detach()).abs().max()) <= float(A.detach().abs().max()) * 2.0 ** -15
            wa, wc, wb = torch.randn_like(A), torch.randn_like(c), torch.randn_like(B)
            total = total + (t.A * wa).sum() + (t.cs * wc).sum() + (t.B * wb).sum()
            ref_total = ref_total + (A * wa).sum() + (s * c * wc).sum() + (B * wb).sum()
            if beta is not None:
                key = {"proj": "attn.proj", "fc1": "mlp.fc1", "fc2": "mlp.fc2"}[which]
                assert torch.allclose(t.bias, (st["blocks.%d.%s.bias" % (l, key)] + s * beta).detach(), atol=1e-6)
                wbias = torch.randn_like(beta)
                total = total + (t.bias * wbias).sum(); ref_total = ref_total + (s * beta * wbias).sum()
    total.backward(); ref_total.backward()
    for k, leaf in leaves.items():
        got = getattr(vit, k).grad
        assert torch.allclose(got, leaf.grad, rtol=1e-4, atol=1e-5), k
    # cache: within one root forward (token) the staged graph is shared; a new forward or a parameter change rebuilds
    vit.__dict__["_cara_fwd_token"] = tok = object()
    first, _ = staging.staged(vit)
    assert staging.staged(vit)[0] is first
    vit.__dict__["_cara_fwd_token"] = object()
    assert staging.staged(vit)[0] is not first
    with torch.no_grad():
        frozen, _ = staging.staged(vit)
        assert staging.staged(vit)[0] is frozen
        vit.CP_R1.add_(1.0)
        assert staging.staged(vit)[0] is not frozen


@pytest.mark.gpu
@pytest.mark.parametrize("L_,C,H,R,sink", [(2, 256, 4, 8, False), (12, 768, 12, 16, True), (3, 1024, 16, 32, True), (2, 1280, 16, 20, False)])
def test_stage_terms_kernel_forward_and_backward_vs_torch(L_, C, H, R, sink):
    """``cara_stage_terms`` (forward and chain rule) against the torch formulation and its autograd, with dense
    gradients and with gradients that are [..., :R] views into Rp-wide buffers (what ops.GradSink hands back)."""
    from cara_b200 import kernels as K
    from cara_b200 import staging
    from tests import _torch_ref
    gen = torch.Generator().manual_seed(3)
    rnd = lambda *s: torch.randn(*s, generator=gen).cuda()                              # noqa: E731
    D, Rp = C // H, K.round_rank(R)
    P = {"CP_A1": rnd(3 * L_, R), "CP_A2": rnd(C, R), "CP_A3": rnd(H, R), "CP_A4": rnd(D, R), "CP_P1": rnd(9 * L_, R),
         "CP_P2": rnd(C, R), "CP_P3": rnd(C, R), "CP_R1": rnd(R), "CP_R2": rnd(R), "CP_bias1": rnd(C), "CP_bias2": rnd(4 * C),
         "CP_bias3": rnd(C)}
    ai, pi, mi = torch.arange(L_).cuda() * 3, torch.arange(L_).cuda() * 9, torch.arange(L_).cuda() * 9 + 1
    s_a, s_m = rnd(L_).abs() + 0.5, rnd(L_).abs() + 0.5
    fb = [rnd(L_, C), rnd(L_, 4 * C), rnd(L_, C)]
    names = ("CP_A1", "CP_A3", "CP_A4", "CP_P1", "CP_P2", "CP_R1", "CP_R2", "CP_bias1", "CP_bias2", "CP_bias3")
    leaf_k = {n: P[n].clone().requires_grad_(True) for n in names}
    leaf_t = {n: P[n].clone().requires_grad_(True) for n in names}
    pads = {"cs_qkv": torch.zeros(L_, 3, Rp).cuda(), "cs_proj": torch.zeros(L_, 1, Rp).cuda(), "cs_fc1": torch.zeros(L_, 4, Rp).cuda(),
            "cs_fc2": torch.zeros(L_, 1, Rp).cuda()}
    const = {"ai": ai.int(), "pi": pi.int(), "mi": mi.int(), "s_a": s_a, "s_m": s_m, "fb_proj": fb[0], "fb_fc1": fb[1],
             "fb_fc2": fb[2], "D": D, "Rp": Rp, "pads": pads}
    got = staging.StageFunction.apply(*[leaf_k[n] for n in names], const)
    Pt = dict(P); Pt.update(leaf_t)
    want = _torch_ref.stage_terms(Pt, ai, pi, mi, s_a, s_m, *fb)
    for name, a, b in zip(staging._STAGED, got, want):
        assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-6, atol=1e-7), name
    for k in pads:
        w = dict(zip(staging._STAGED, want))[k]
        assert torch.allclose(pads[k][..., :R], w, rtol=1e-6, atol=1e-7) and float(pads[k][..., R:].abs().max() if Rp > R else 0.0) == 0.0
    grads = []
    for a in want:
        g = rnd(*a.shape)
        if sink and a.shape[-1] == R and a.dim() >= 2 and not (a.dim() == 2 and a.shape[0] == L_ and a.shape[1] in (C, 4 * C)):
            buf = torch.zeros(*a.shape[:-1], Rp).cuda()
            buf[..., :R] = g
            grads.append((buf[..., :R], g))
        else:
            grads.append((g, g))
    torch.autograd.backward(list(got), [g[0] for g in grads])
    torch.autograd.backward(list(want), [g[1] for g in grads])
    for n in names:
        assert torch.allclose(leaf_k[n].grad, leaf_t[n].grad, rtol=2e-4, atol=1e-5), (n, float((leaf_k[n].grad - leaf_t[n].grad).abs().max()))


def test_state_dict_schema_roundtrip():
    vit, st, g = _small()
    sd = vit.state_dict()
    assert set(sd) == set(st)
    for k in st:
        assert tuple(sd[k].shape) == tuple(st[k].shape), k
    assert set(O.cp_shapes(g)) | set(O.backbone_shapes(g)) == set(sd)


def test_kernel_wrappers_refuse_cpu_tensors():
    from cara_b200 import kernels as K
    from cara_b200._lib import CaraLibraryError
    with pytest.raises(CaraLibraryError):
        K.gemm_cp(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(256, 64, dtype=torch.bfloat16))
    with pytest.raises(CaraLibraryError):
        K.ln_fwd(torch.zeros(8, 128), torch.ones(128), torch.zeros(128))
    with pytest.raises(CaraLibraryError):
        K.factor_operands(torch.zeros(8, 4), 16)


def test_vtab_config_and_cli_surface():
    import importlib.util
    spec = importlib.util.spec_from_file_location("vtab_config", os.path.join(ROOT, "image_classification", "vtab_config.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    assert len(mod.config) == 19
    for name, c in mod.config.items():
        assert set(c) == {"init_mean", "init_std", "scale", "seed", "logger"}, name


def test_npz_checkpoint_import_roundtrip(tmp_path):
    """vit_cp.py:155 loads ./ViT-B_16.npz through timm's JAX importer: write a Flax-layout .npz from a random
    model (inverse mapping, independent code) and check create_model(checkpoint_path=...) reproduces every
    parameter, including a position-embedding grid resample and the skipped head on a class-count mismatch."""
    import numpy as np
    import torch
    from cara_b200.vit import create_model
    torch.manual_seed(3)
    C, H, L = 64, 4, 2
    src = create_model("vit_base_patch16_224_in21k", embed_dim=C, depth=L, num_heads=H, num_classes=7, img_size=32)
    with torch.no_grad():
        for p in src.parameters():
            p.copy_(torch.randn_like(p))
    sd = {k: v.detach().numpy() for k, v in src.state_dict().items()}
    z = {"embedding/kernel": sd["patch_embed.proj.weight"].transpose(2, 3, 1, 0), "embedding/bias": sd["patch_embed.proj.bias"],
         "cls": sd["cls_token"], "Transformer/posembed_input/pos_embedding": sd["pos_embed"],
         "Transformer/encoder_norm/scale": sd["norm.weight"], "Transformer/encoder_norm/bias": sd["norm.bias"],
         "head/kernel": sd["head.weight"].T, "head/bias": sd["head.bias"]}
    D = C // H
    for i in range(L):
        b, m = "Transformer/encoderblock_%d/" % i, "Transformer/encoderblock_%d/MultiHeadDotProductAttention_1/" % i
        t = "blocks.%d." % i
        z[b + "LayerNorm_0/scale"], z[b + "LayerNorm_0/bias"] = sd[t + "norm1.weight"], sd[t + "norm1.bias"]
        z[b + "LayerNorm_2/scale"], z[b + "LayerNorm_2/bias"] = sd[t + "norm2.weight"], sd[t + "norm2.bias"]
        for j, n in enumerate(("query", "key", "value")):
            z[m + n + "/kernel"] = sd[t + "attn.qkv.weight"][j * C:(j + 1) * C].T.reshape(C, H, D)
            z[m + n + "/bias"] = sd[t + "attn.qkv.bias"][j * C:(j + 1) * C].reshape(H, D)
        z[m + "out/kernel"] = sd[t + "attn.proj.weight"].T.reshape(H, D, C)
        z[m + "out/bias"] = sd[t + "attn.proj.bias"]
        for r in range(2):
            z[b + "MlpBlock_3/Dense_%d/kernel" % r] = sd[t + "mlp.fc%d.weight" % (r + 1)].T
            z[b + "MlpBlock_3/Dense_%d/bias" % r] = sd[t + "mlp.fc%d.bias" % (r + 1)]
    path = str(tmp_path / "ViT-tiny.npz")
    np.savez(path, **z)
    dst = create_model("vit_base_patch16_224_in21k", checkpoint_path=path, embed_dim=C, depth=L, num_heads=H,
                       num_classes=7, img_size=32)
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    # other class count: head keeps its own init; other image size: grid rows are resampled, class token row kept
    dst2 = create_model("vit_base_patch16_224_in21k", checkpoint_path=path, embed_dim=C, depth=L, num_heads=H,
                        num_classes=5, img_size=64)
    assert dst2.head.weight.shape == (5, C) and dst2.pos_embed.shape == (1, 17, C)
    assert torch.equal(dst2.pos_embed[:, 0], src.pos_embed[:, 0])
    assert torch.equal(dst2.blocks[1].mlp.fc2.weight, src.blocks[1].mlp.fc2.weight)


def test_cosine_lr_matches_timm_closed_form():
    """cara_b200.train.cosine_lr against the closed form of timm 0.4.12's CosineLRScheduler._get_lr for the arguments
    of vit_cp.py:187 (t_initial=100, warmup_t=10, lr_min=1e-5, warmup_lr_init=1e-6, decay_rate=0.1; t_mul 1, no warmup
    prefix, cycle_limit 0), written out independently here."""
    import math
    from cara_b200 import train as T

    def timm_lr(t, base=1e-3, t_initial=100, warmup_t=10, lr_min=1e-5, warmup_lr_init=1e-6, decay_rate=0.1):
        if t < warmup_t:
            return warmup_lr_init + t * ((base - warmup_lr_init) / warmup_t)
        i = t // t_initial
        t_curr = t - t_initial * i
        gamma = decay_rate ** i
        lo, hi = lr_min * gamma, base * gamma
        return lo + 0.5 * (hi - lo) * (1 + math.cos(math.pi * t_curr / t_initial))

    for base in (1e-3, 5e-4):
        for t in range(0, 230):
            assert T.cosine_lr(t, base_lr=base) == pytest.approx(timm_lr(t, base=base), rel=1e-12, abs=0), (t, base)
    assert T.cosine_lr(0) == 1e-6 and T.cosine_lr(10) == pytest.approx(1e-5 + 0.5 * (1e-3 - 1e-5) * (1 + math.cos(0.1 * math.pi)))
    assert T.cosine_lr(9) == pytest.approx(1e-6 + 9 * (1e-3 - 1e-6) / 10)


def test_per_step_lr_sequence_of_the_reference_loop():
    """The learning rate of EVERY optimizer step of reference vit_cp.py:26-59: the scheduler leaves warmup_lr_init in the
    optimizer at construction and is stepped AFTER opt.step(), so batch 0 of epoch e runs at the rate of epoch e-1; it is
    dropped at the periodic test of epoch 50.  Simulated here with the reference's control flow and compared with
    cara_b200.train.EpochCosineSchedule as image_classification/vit_cp.py drives it."""
    from cara_b200 import train as T
    batches = 3
    # the reference's control flow, literally
    want, lr, sched_on = [], T.cosine_lr(0), True
    for epoch in range(100):
        for _ in range(batches):
            want.append(lr)                              # opt.step() uses the current rate
            if sched_on:
                lr = T.cosine_lr(epoch)                  # sched.step(epoch)
        if epoch % 10 == 0 and epoch != 0 and epoch >= 50:
            sched_on = False
    got, s = [], T.EpochCosineSchedule(base_lr=1e-3)
    cur = s.lr
    for epoch in range(100):
        for _ in range(batches):
            got.append(cur)
            cur = s.after_step(epoch)
        if epoch % 10 == 0 and epoch != 0:
            s.after_test(epoch)
    assert got == want
    assert got[0] == got[batches - 1] == 1e-6                     # all of epoch 0 at warmup_lr_init
    assert got[batches] == 1e-6 and got[batches + 1] == pytest.approx(1e-6 + (1e-3 - 1e-6) / 10)
    assert got[-1] == T.cosine_lr(50) and got[51 * batches + 1] == T.cosine_lr(50)   # frozen after epoch 50's test


def test_create_model_raises_on_missing_checkpoint(tmp_path):
    from cara_b200.vit import create_model
    missing = os.path.join(str(tmp_path), "ViT-B_16.npz")
    with pytest.raises(FileNotFoundError):
        create_model("vit_base_patch16_224_in21k", checkpoint_path=missing, depth=1, embed_dim=128, num_heads=2)
    with pytest.warns(UserWarning):
        create_model("vit_base_patch16_224_in21k", checkpoint_path=missing, allow_missing_checkpoint=True, depth=1,
                     embed_dim=128, num_heads=2)


def test_hand_over_links_between_autograd_functions():
    """The Python side of the fused hand-overs (cara_b200.ops): a backward RowsLink is armed only when the projection
    trains through the gradient sink with one slice, is consumed exactly once, and the per-block link decision refuses
    models that cannot use the fused LayerNorm kernels (no CUDA calls here)."""
    from cara_b200 import ops, vit
    from src.cara.cara import cara

    class FakeOps:
        slices, rp = 1, 16

    class FakeSink:
        def __init__(self):
            self.t = {"dcs1": torch.zeros(3, 1, 16), "dcs3": torch.zeros(3, 1, 16)}

    T = torch.zeros(5, 16)
    sink = (FakeSink(), 1, 2)
    link = ops.RowsLink()
    ops._arm_backward_link(link, FakeOps(), T, sink, True)
    assert link.T is T and link.ops is not None and link.dcs.shape == (1, 16)
    assert link.dcs.data_ptr() == sink[0].t["dcs1"][2].data_ptr()          # the LayerNorm accumulates in place
    for bad in [dict(train=False), dict(sink=None), dict(T=None)]:
        l2 = ops.RowsLink()
        kw = dict(ops=FakeOps(), T=T, sink=sink, train=True)
        kw.update(bad)
        ops._arm_backward_link(l2, kw["ops"], kw["T"], kw["sink"], kw["train"])
        assert l2.T is None and l2.ops is None
    wide = FakeOps(); wide.slices = 3                                       # qkv / fc1 never get a backward link
    l3 = ops.RowsLink()
    ops._arm_backward_link(l3, wide, T, sink, True)
    assert l3.T is None
    assert ops._take_dT(None) is None and ops._take_dT(link) is None        # the LayerNorm backward never ran
    link.dT = torch.ones(5, 48)
    got = ops._take_dT(link)
    assert got is not None and link.dT is None and link.T is None and link.dcs is None and ops._take_dT(link) is None

    # per-block decision: un-adapted model -> no links; fp32 parity mode -> no links; merged (forward un-patched) -> none
    m = vit.create_model("vit_base_patch16_224_in21k", depth=1)
    assert m._rows_links(m.blocks[0]) is None
    m = cara({"model": m, "rank": 8, "scale": 1.0, "l_mu": 1.0, "l_std": 0.0})
    m.__dict__["cara_precision"] = "fp32"
    assert m._rows_links(m.blocks[0]) is None
    m.__dict__["cara_precision"] = "bf16"
    m.blocks[0].attn.__dict__.pop("forward")
    assert m._rows_links(m.blocks[0]) is None
