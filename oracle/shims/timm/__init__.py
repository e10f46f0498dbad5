"""Test-only stand-in for the parts of ``timm==0.4.12`` the CaRA reference touches.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference pins
timm==0.4.12 (pyproject.toml:12) but the package is not installed in this
image and there is no network, so the handful of classes the reference
dispatches on (cara.py:110,147,157) and the factory it calls
(vit_cp.py:155, tests/test_cara.py:19) are restated here from the published
timm 0.4.12 behaviour.  Nothing in the product path imports this package.
"""
from . import models  # noqa: F401

__version__ = "0.4.12+oracle.shim"
