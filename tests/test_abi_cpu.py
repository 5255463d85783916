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
    """The oracle is test infrastructure: nothing under the product package, scripts/ (nor the GPU arm of bench.py) may import it."""
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "quantised-bayesian-nets_b200")
    walks = list(os.walk(pkg)) + list(os.walk(os.path.join(root, "scripts")))      # measurement scripts count as product side
    for dirpath, _, files in walks:
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


def test_argument_validation_needs_no_gpu(lib):
    """Error behaviour of the boundary (include/qbn.h:8-12): invalid arguments are rejected up front with
    QBN_ERR_INVALID_ARG and a message naming the function — before any CUDA call, so this runs without a device."""
    l = lib.load()
    buf = ctypes.create_string_buffer(64)               # a non-null pointer that is never dereferenced
    p = ctypes.cast(buf, ctypes.c_void_p)
    cases = {
        "qbn_i8_add": lambda: l.qbn_i8_add(p, 0.1, 0, p, 0.1, 0, 0, -1, 0.1, 0, 0, 255, p, None),                  # n == 0
        "qbn_i8_relu": lambda: l.qbn_i8_relu(p, 16, 0, 200, 100, p, None),                                          # lo > hi
        "qbn_i8_avgpool": lambda: l.qbn_i8_avgpool(p, 1, 6, 6, 8, 4, 0, 0, 255, p, None),                           # 4 does not divide 6
        "qbn_quantize_u8": lambda: l.qbn_quantize_u8(p, 16, 0.0, 0, 0, 255, p, None),                               # scale == 0
        "qbn_dequantize_u8": lambda: l.qbn_dequantize_u8(None, 16, 0.1, 0, p, None),                                # null input
        "qbn_softmax_accumulate": lambda: l.qbn_softmax_accumulate(p, 4, 8, 129, p, 0, None),                       # K > 128
        "qbn_cls_metrics": lambda: l.qbn_cls_metrics(p, p, 8, 10, 1.0, 64, p, None),                                # n_bins > 32
        "qbn_maxpool2x2": lambda: l.qbn_maxpool2x2(p, 1, 1, 8, 4, p, None),                                         # H == 1
        "qbn_nchw_to_nhwc": lambda: l.qbn_nchw_to_nhwc(p, 70000, 3, 16, p, None),                                   # B > 65535
        "qbn_avgpool_all": lambda: l.qbn_avgpool_all(p, 0, 16, 8, 0.0, p, None),                                    # B == 0
    }
    for name, call in cases.items():
        assert call() == -1, name
        msg = l.qbn_last_error()
        msg = msg.decode() if isinstance(msg, bytes) else msg
        assert name in msg and "invalid argument" in msg, (name, msg)
