"""The reference's own module-setup tests (reference tests/test_cara.py:43-90) against the drop-in module.

``create_model`` comes from cara_b200.vit instead of timm (timm is not a dependency); everything else --
config keys, fixture seeding, assertions -- is as in the reference.  The reference's fifth test (forward shape)
needs the GPU and lives in tests/test_parity_gpu.py::test_reference_forward_shape.
"""
import random
from typing import Any, Dict

import numpy as np
import pytest
import torch as th

from cara_b200.vit import create_model
from src.cara.cara import cara


def _get_vit() -> th.nn.Module:
    return create_model("vit_base_patch16_224_in21k", drop_path_rate=0.1)


def _get_cara_config() -> Dict[str, Any]:
    random.seed(0)
    th.manual_seed(0)
    np.random.seed(0)
    th.cuda.manual_seed_all(0)
    th.backends.cudnn.deterministic = True
    th.backends.cudnn.benchmark = False
    return {"model": _get_vit(), "rank": 32, "scale": 1.0, "l_mu": 1.0, "l_std": 0.0}


def test_vit_without_cara():
    vit = _get_vit()
    for n in ("CP_A1", "CP_A2", "CP_A3", "CP_A4", "CP_P1", "CP_P2", "CP_P3", "CP_R1", "CP_R2"):
        assert not hasattr(vit, n)


def test_vit_with_cara():
    vit = cara(_get_cara_config())
    for n in ("CP_A1", "CP_A2", "CP_A3", "CP_A4", "CP_P1", "CP_P2", "CP_P3", "CP_R1", "CP_R2"):
        assert hasattr(vit, n)


def test_cara_zero_init():
    vit = cara(_get_cara_config())
    assert th.allclose(vit.CP_A2, th.zeros_like(vit.CP_A2))
    assert th.allclose(vit.CP_P2, th.zeros_like(vit.CP_P2))


def test_cara_lambda_init():
    vit = cara(_get_cara_config())
    assert th.allclose(vit.CP_R1, th.ones_like(vit.CP_R1))
    assert th.allclose(vit.CP_R2, th.ones_like(vit.CP_R2))


def test_cara_surface_matches_reference():
    """Shapes (cara.py:112-125), row maps, child attributes (cara.py:148-162), same instance back."""
    cfg = _get_cara_config()
    vit = cara(cfg)
    assert vit is cfg["model"]
    import src.cara.cara as mod
    assert mod.global_model is vit
    shapes = {n: tuple(getattr(vit, n).shape) for n in ("CP_A1", "CP_A2", "CP_A3", "CP_A4", "CP_P1", "CP_P2",
                                                         "CP_P3", "CP_R1", "CP_R2", "CP_bias1", "CP_bias2", "CP_bias3")}
    assert shapes == {"CP_A1": (36, 32), "CP_A2": (768, 32), "CP_A3": (12, 32), "CP_A4": (64, 32),
                      "CP_P1": (108, 32), "CP_P2": (768, 32), "CP_P3": (768, 32), "CP_R1": (32,), "CP_R2": (32,),
                      "CP_bias1": (768,), "CP_bias2": (3072,), "CP_bias3": (768,)}
    for l, blk in enumerate(vit.blocks):
        assert (blk.attn.attn_idx, blk.attn.idx, blk.mlp.idx) == (3 * l, 9 * l, 9 * l + 1)
        for m in (blk.attn, blk.mlp):
            assert isinstance(m.dp, th.nn.Dropout) and m.dp.p == 0.1 and m.s == 1.0 and m.dim == 32
        assert blk.attn.forward.__func__ is mod.cp_attn and blk.mlp.forward.__func__ is mod.cp_mlp
    assert len(vit.state_dict()) == 164
    vit.reset_classifier(100)
    names = [n for n, _ in vit.named_parameters() if "CP" in n or "head" in n]
    assert len(names) == 14


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU instead of silently computing on the CPU."""
    from cara_b200._lib import CaraLibraryError
    vit = cara(_get_cara_config())
    with pytest.raises(CaraLibraryError):
        vit(th.randn(1, 3, 224, 224))
    with pytest.raises(CaraLibraryError):
        vit.blocks[0].attn(th.randn(1, 197, 768))
