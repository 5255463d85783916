"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/qbn.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from qbn_b200 import _lib
    return _lib


def test_exports_match_header(lib):
    header = open(os.path.join(ROOT, "include", "qbn.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(qbn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(handle, n)]
    assert not missing, "library does not export: %s" % missing
    assert declared == set(lib.EXPORTED_SYMBOLS), (declared ^ set(lib.EXPORTED_SYMBOLS))


def test_version_and_no_gpu_error(lib):
    l = lib.load()
    assert l.qbn_version() >= 100
    import torch
    if not torch.cuda.is_available():
        sm = ctypes.c_int()
        assert l.qbn_device_info(ctypes.byref(sm), None, None) != 0  # fails loudly, no CPU fallback
        assert l.qbn_last_error()


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from qbn_b200 import ops
    with pytest.raises(lib.QbnError):
        ops.weight_prep(torch.zeros(2, 2), torch.zeros(2, 2))


def test_sass_has_tcgen05(lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib.LIB_PATH], stdout=subprocess.PIPE).stdout.decode()
    assert "UTCHMMA" in sass and "UTCIMMA" in sass and "LDTM" in sass


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package (nor the GPU arm of bench.py) may import it."""
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "quantised-bayesian-nets_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(dirpath, f)).read())
                for node in ast.walk(tree):
                    names = [a.name for a in node.names] if isinstance(node, ast.Import) else ([node.module or ""] if isinstance(node, ast.ImportFrom) else [])
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (f, names)
    # bench.py: oracle imports only inside the CPU-baseline function
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        imports = [a.name for n in ast.walk(fn) if isinstance(n, ast.Import) for a in n.names]
        if any(i.startswith("oracle") for i in imports):
            assert "cpu" in fn.name.lower() or "reference" in fn.name.lower(), fn.name
