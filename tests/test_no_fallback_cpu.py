"""CPU: the product path has no fallback -- a missing C-ABI library is a loud ImportError at first use, and the
product package never imports oracle/."""
import ast
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_missing_library_raises(monkeypatch):
    from visionllm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "visionllm_b200", "lib", "does_not_exist.so"))
    with pytest.raises(ImportError, match="no CPU/eager fallback"):
        _lib.lib()


def test_product_package_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, "visionllm_b200", "*.py")):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path


def test_cpu_tensors_are_rejected_not_computed():
    import torch
    from visionllm_b200 import ops
    x = torch.randn(4, 64).bfloat16()
    w = torch.randn(8, 64).bfloat16()
    with pytest.raises(RuntimeError):
        ops.linear(x, w)
    with pytest.raises(RuntimeError):
        ops.attention(x.view(1, 4, 1, 64), x.view(1, 4, 1, 64), x.view(1, 4, 1, 64))
