"""Loader for the UNMODIFIED reference (byeonghu-na/INDM) as a CPU oracle.

TEST INFRASTRUCTURE ONLY.  Nothing in `indm_b200/` may import this module.
It is used by `tests/golden/make_golden.py` (run in the build container, where
`/root/reference` exists) to produce the committed golden fixtures, and by the
optional `-m "not gpu"` tests that re-check the restatement in `oracle/` against
the live reference when it is present.  `/root/reference` does not exist on the
GPU box, so nothing on the `-m gpu` / smoke / bench paths touches this file.

What it does: puts four import stubs ahead of the reference on
`sys.path` (`overrides`, `ml_collections`, `tensorflow.io.gfile`, `flowpp_models`),
injects a `torch._six` module (removed from modern torch, needed by
`flow_models/wolf/utils.py:5`), neuters `torch.utils.cpp_extension.load` so that
`import op` does not JIT-compile the reference's CUDA ops (their Python wrappers
route CPU tensors to `upfirdn2d_native` / `F.leaky_relu`, `op/upfirdn2d.py:146-149`,
`op/fused_act.py:87-94`), and runs with cwd = the reference root because
`flow.model_config` is a relative path (`flow_models/flow_model.py:102`).
None of this changes arithmetic.
"""
import contextlib
import importlib
import math
import os
import sys
import types

REF_ROOT = os.environ.get("INDM_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stubs")

# top-level module names the reference uses that collide with nothing of ours
_REF_TOPLEVEL = ("op", "models", "sde_lib", "sampling", "losses", "likelihood",
                 "flow_models", "configs", "utils", "datasets")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


@contextlib.contextmanager
def reference_cwd():
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def _install():
    import torch
    import torch.utils.cpp_extension as cpp_ext

    sys.dont_write_bytecode = True  # the reference tree is read-only
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.inf = math.inf
        six.string_classes = (str,)
        sys.modules["torch._six"] = six
        torch._six = six
    if not getattr(cpp_ext.load, "_indm_oracle_stub", False):
        def _no_jit(*a, **k):
            return None
        _no_jit._indm_oracle_stub = True
        cpp_ext.load = _no_jit
    for p in (REF_ROOT, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _STUBS)


def load(*names):
    """Import reference modules by their top-level names, e.g. load('sde_lib', 'models.ncsnpp')."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install()
    with reference_cwd():
        mods = [importlib.import_module(n) for n in names]
    return mods[0] if len(mods) == 1 else mods


def get_config(path: str):
    """Load a reference config file, e.g. 'configs/vp/CIFAR10/indm_fid.py', device forced to CPU."""
    import importlib.util
    import torch
    _install()
    with reference_cwd():
        spec = importlib.util.spec_from_file_location("_indm_ref_cfg_" + path.replace("/", "_"),
                                                      os.path.join(REF_ROOT, path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        cfg = mod.get_config()
    cfg.device = torch.device("cpu")
    return cfg
