"""Test-only stand-in for the two ``tensorly==0.8.1`` calls the CaRA reference makes.

TEST INFRASTRUCTURE ONLY.  The reference pins tensorly==0.8.1
(pyproject.toml:11) and calls ``tl.set_backend("pytorch")`` (cara.py:10) and
``tl.cp_to_tensor((weights, factors))`` (cara.py:27,52,76,88).  tensorly's
published CP reconstruction is: scale the first factor's columns by the
weights, multiply by the transposed Khatri-Rao product of the remaining
factors (row index in C order, first remaining factor slowest), and fold the
[I0, I1*...*In] matrix back to shape (I0, I1, ..., In).
"""
import torch

_backend = "numpy"


def set_backend(name):
    global _backend
    _backend = name


def get_backend():
    return _backend


def khatri_rao(matrices):
    rank = matrices[0].shape[1]
    out = matrices[0]
    for m in matrices[1:]:
        if m.shape[1] != rank:
            raise ValueError("All matrices must have the same number of columns")
        out = (out.unsqueeze(1) * m.unsqueeze(0)).reshape(-1, rank)
    return out


def cp_to_tensor(cp_tensor, mask=None):
    weights, factors = cp_tensor
    rank = factors[0].shape[1]
    for f in factors:
        if f.ndim != 2 or f.shape[1] != rank:
            raise ValueError("All the factors of a CP tensor should have the same number of column")
    shape = [f.shape[0] for f in factors]
    lead = factors[0] if weights is None else factors[0] * weights
    if len(factors) == 1:
        return lead.sum(dim=1)
    flat = lead @ khatri_rao(list(factors[1:])).T
    return flat.reshape(shape)
