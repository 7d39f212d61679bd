"""Guards of the tier rules: the product package never imports the oracle (or the reference tree), the oracle is only reachable from
the checker legs of bench.py / __graft_entry__.smoke(), and the product fails loudly without its CUDA library / on CPU tensors."""
import ast
import glob
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name, node.lineno
        elif isinstance(node, ast.ImportFrom):
            yield (node.module or ""), node.lineno


def test_product_package_never_imports_oracle_or_reference():
    bad = []
    for path in glob.glob(os.path.join(ROOT, "indm_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        for mod, line in _imports(path):
            if mod == "oracle" or mod.startswith("oracle."):
                bad.append((os.path.relpath(path, ROOT), line, mod))
        if "/root/reference" in src:
            bad.append((os.path.relpath(path, ROOT), 0, "/root/reference"))
    assert not bad, bad


def test_bench_reaches_the_oracle_only_from_its_cpu_legs():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_sample", "cpu_train_sample", "_reference_sampler"}     # _reference_sampler: the `--impl reference` arm and the cpu_baseline leg only
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = [m for node in ast.walk(fn) if isinstance(node, (ast.Import, ast.ImportFrom))
                for m in ([a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]) if m.split(".")[0] == "oracle"]
        assert not uses or fn.name in allowed, (fn.name, uses)
    top = [m for node in tree.body if isinstance(node, (ast.Import, ast.ImportFrom))
           for m in ([a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""])]
    assert not any(m.split(".")[0] == "oracle" for m in top)
    assert "/root/reference" not in open(os.path.join(ROOT, "bench.py")).read()


def test_gpu_tests_and_smoke_do_not_read_the_reference_tree():
    for path in glob.glob(os.path.join(ROOT, "tests", "test_*_gpu.py")) + [os.path.join(ROOT, "__graft_entry__.py")]:
        src = open(path).read()
        assert "/root/reference" not in src and "ref_loader" not in src, path


def test_missing_library_is_a_loud_error(monkeypatch, tmp_path):
    from indm_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libindm_b200.so"))
    monkeypatch.setattr(_lib, "_handle", None, raising=False)
    for name in ("_LIB", "_lib", "_h"):
        if hasattr(_lib, name):
            monkeypatch.setattr(_lib, name, None)
    with pytest.raises((RuntimeError, OSError)):
        _lib.lib()


def test_ops_refuse_cpu_tensors():
    from indm_b200 import op
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(torch.zeros(1, 2, 4, 4), torch.zeros(2))
