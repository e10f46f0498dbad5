"""Pin the CPU oracle (oracle/cara_oracle.py) against outputs of the reference itself.

The golden files were produced by tests/golden/make_golden.py from the UNMODIFIED
/root/reference/src/cara/cara.py (reference tests hold no numeric vectors).
"""
import os

import numpy as np
import pytest
import torch

from oracle import cara_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def test_halves_match_reference_fp64():
    z = np.load(os.path.join(G, "ref_halves_fp64.npz"))
    g = O.Geometry(depth=2, rank=8, num_classes=10)
    st = O.synthetic_state(g, dtype=torch.float64)
    x = torch.from_numpy(z["x"])
    layer, s = int(z["layer"]), float(z["scale"])
    assert (int(z["attn_idx"]), int(z["idx"]), int(z["mlp_idx"])) == (3 * layer, 9 * layer, 9 * layer + 1)
    assert _rel(O.attn_half(st, g, layer, x, s), z["attn"]) < 1e-13
    assert _rel(O.mlp_half(st, g, layer, x, s), z["mlp"]) < 1e-13


@pytest.mark.parametrize("fname,geom,dtype,tol", [
    ("ref_vitb_d2_r8_fp64.npz", dict(depth=2, rank=8, num_classes=10), torch.float64, 1e-11),
    ("ref_vitb_d12_r16_fp32.npz", dict(depth=12, rank=16, num_classes=100), torch.float32, 2e-4),
])
def test_model_fwd_bwd_match_reference(fname, geom, dtype, tol):
    z = np.load(os.path.join(G, fname))
    g = O.Geometry(**geom)
    st = O.synthetic_state(g, dtype=dtype)
    x, y = O.synthetic_batch(g, int(z["batch"]), dtype=dtype)
    logits, loss, grads = O.loss_and_grads(st, g, x, y, float(z["scale"]))
    assert _rel(logits, z["logits"]) < tol
    assert abs(float(loss) - float(z["loss"])) < tol * 10
    for k, gr in grads.items():
        ref = z["grad." + k]
        assert gr.shape == ref.shape, k
        assert _rel(gr, ref) < tol * 20, (k, _rel(gr, ref))


def test_factored_chain_equals_materialised_delta():
    """the (A, c, B, beta) factoring the CUDA kernels use reproduces the
    reference's materialised delta-weights (fp64)."""
    g = O.Geometry(embed_dim=64, depth=2, num_heads=4, rank=8, num_classes=7, img=32, patch=16)
    st = O.synthetic_state(g, dtype=torch.float64)
    s, l = 0.7, 1
    merged = O.merged_weights(st, g, s)
    for which, wkey in (("qkv", "attn.qkv"), ("proj", "attn.proj"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2")):
        A, c, B, beta = O.adapter_terms(st, g, l, which)
        W = st["blocks.%d.%s.weight" % (l, wkey)]
        if which == "fc2":
            dW = (B * c[0]) @ A.T
        else:
            dW = torch.cat([(B * ck) @ A.T for ck in c], dim=0)
        assert _rel(W + s * dW, merged["blocks.%d.%s.weight" % (l, wkey)]) < 1e-13
    x, _ = O.synthetic_batch(g, 3, dtype=torch.float64)
    assert _rel(O.forward_plain(merged, g, x), O.forward(st, g, x, s)) < 1e-12


def test_default_init_is_identity_and_shapes():
    g = O.Geometry(embed_dim=64, depth=2, num_heads=4, rank=8, num_classes=7, img=32, patch=16)
    st = O.synthetic_state(g, dtype=torch.float64)
    cp = O.declare_cp(g, dtype=torch.float64)
    assert {k: tuple(v.shape) for k, v in cp.items()} == O.cp_shapes(g)
    assert float(cp["CP_A2"].abs().max()) == 0.0 and float(cp["CP_P2"].abs().max()) == 0.0   # tests/test_cara.py:79-83
    assert torch.equal(cp["CP_R1"], torch.ones(8, dtype=torch.float64))                       # tests/test_cara.py:86-90
    st.update(cp)
    x, _ = O.synthetic_batch(g, 2, dtype=torch.float64)
    plain = {k: v for k, v in st.items() if not k.startswith("CP_")}
    assert _rel(O.forward(st, g, x, 3.0), O.forward_plain(plain, g, x)) < 1e-13


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(50, dtype=torch.float64); ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-4)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(50, dtype=torch.float64)
        ref.grad = gr.clone(); opt.step()
        p, m, v = O.adamw_update(p, gr, m, v, step)
        assert _rel(p, ref.detach()) < 1e-13
