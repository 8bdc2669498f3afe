"""ref_import -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference modules (`networks`, `smr_utils`) from
/root/reference with (a) `kaolin` replaced by oracle/kaolin_shim.py and (b) the
absent, off-path dependencies (`timm`, `pytorch3d`) stubbed.  Only usable in the
build container: /root/reference does not exist on the GPU box, so nothing that
runs there (gpu tests, smoke, bench) may call this -- it is used by
tests/golden/make_golden.py to produce committed fixtures and by the CPU tests
that re-check those fixtures when the reference tree is present.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "networks.py"))


def import_reference():
    """Returns (networks_module, smr_utils_module) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import kaolin_shim
    kaolin_shim.install()
    # off-path dependencies that are not installed here
    if "timm" not in sys.modules:
        try:
            import timm  # noqa: F401
        except Exception:
            sys.modules["timm"] = types.ModuleType("timm")
    if "pytorch3d" not in sys.modules:
        try:
            import pytorch3d.loss  # noqa: F401
        except Exception:
            p3d = types.ModuleType("pytorch3d")
            loss = types.ModuleType("pytorch3d.loss")

            def chamfer_distance(*a, **k):
                raise NotImplementedError("pytorch3d is not installed (off the hot path)")
            loss.chamfer_distance = chamfer_distance
            p3d.loss = loss
            sys.modules["pytorch3d"] = p3d
            sys.modules["pytorch3d.loss"] = loss
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # networks.py:252 hard-codes `.cuda()`; make it a no-op on CPU-only hosts.
    if not torch.cuda.is_available() and not getattr(torch.Tensor.cuda, "__mm_noop__", False):
        def _cuda_noop(self, *a, **k):
            return self
        _cuda_noop.__mm_noop__ = True
        torch.Tensor.cuda = _cuda_noop
    import networks as ref_networks
    import smr_utils as ref_smr_utils
    return ref_networks, ref_smr_utils
