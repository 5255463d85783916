"""cuBLAS GEMM throughput of this box at 8192^3 in TF32 and INT8 (SURVEY 8d asks for them beside the driver's bf16 figure in
MEASURED_PEAKS.json): the tensor-side roofline denominators of the kind::tf32 / kind::i8 contractions.  Library GEMMs, used only as a
yardstick.  Usage: python scripts/measure_peaks.py > profiles/r02_measured_peaks_tf32_int8.json"""
import json

import torch


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    n = 8192
    flop = 2.0 * n ** 3
    out = {"n": n, "device": torch.cuda.get_device_name(0)}
    a, b = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    out["tf32_tflops"] = flop / timeit(lambda: a @ b) / 1e12
    torch.backends.cuda.matmul.allow_tf32 = False
    out["fp32_tflops"] = flop / timeit(lambda: a @ b, 5) / 1e12
    ab, bb = a.bfloat16(), b.bfloat16()
    out["bf16_tflops"] = flop / timeit(lambda: ab @ bb) / 1e12
    ai = torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8)
    bi = torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8)
    try:
        out["int8_tops"] = flop / timeit(lambda: torch._int_mm(ai, bi)) / 1e12
    except Exception as e:          # noqa: BLE001
        out["int8_tops"] = None
        out["int8_error"] = str(e)[:200]
    # long run (sustained clocks): 2 s of TF32 GEMMs
    reps = max(20, int(2.0 / (flop / (out["tf32_tflops"] * 1e12))))
    torch.backends.cuda.matmul.allow_tf32 = True
    out["tf32_tflops_sustained"] = flop / timeit(lambda: a @ b, reps) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
