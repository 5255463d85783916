"""Process-wide switches of the hot path."""
from ._lib import QBN_MATH_FP32, QBN_MATH_TF32

_cfg = {"math_mode": QBN_MATH_FP32, "pdl": False}


def set_math_mode(mode):
    """'fp32' (CUDA-core FFMA, rtol 1e-5 parity) or 'tf32' (tcgen05 kind::tf32, rtol 1e-3)."""
    _cfg["math_mode"] = {"fp32": QBN_MATH_FP32, "tf32": QBN_MATH_TF32}[mode] if isinstance(mode, str) else int(mode)


def math_mode():
    return _cfg["math_mode"]


def tf32_eligible(C, N, lrt):
    """Shapes the tcgen05 kernels take: 16-byte K-chunks need C % 4 == 0; one CTA holds the full
    N extent in TMEM (two accumulators for LRT)."""
    n_pad = (N + 15) // 16 * 16
    return C % 4 == 0 and (2 * n_pad <= 512 if lrt else n_pad <= 256)


def pick_math_mode(C, N, lrt):
    m = math_mode()
    if m == QBN_MATH_TF32 and not tf32_eligible(C, N, lrt):
        return QBN_MATH_FP32  # still a libqbn CUDA kernel (the FFMA one), never a CPU/PyTorch path
    return m


def set_pdl(enabled):
    """Programmatic dependent launch inside the engines' CUDA graphs (include/qbn.h: qbn_set_pdl): the prologue of every planar
    conv launch overlaps the tail of the previous kernel."""
    _cfg["pdl"] = bool(enabled)


def pdl():
    return _cfg["pdl"]
