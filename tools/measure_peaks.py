#!/usr/bin/env python
"""The denominators of the rooflines, measured on this GPU: HBM copy, dense BF16 and TF32 GEMM (cuBLAS,
8192^3) and fp32 FFMA (own kernel, s2d_debug_ffma).  Best of N with CUDA events.  usage: measure_peaks.py > profiles/..."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import _lib  # noqa: E402


def best(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    b = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        b = ms if b is None else min(b, ms)
    return b


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    n = 8192
    a16, b16 = torch.randn(n, n, device="cuda", dtype=torch.bfloat16), torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    out["bf16_gemm_tflops"] = 2.0 * n ** 3 / (best(lambda: torch.matmul(a16, b16)) * 1e-3) / 1e12
    a32, b32 = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    out["tf32_gemm_tflops"] = 2.0 * n ** 3 / (best(lambda: torch.matmul(a32, b32)) * 1e-3) / 1e12
    torch.backends.cuda.matmul.allow_tf32 = False
    out["fp32_gemm_tflops_cublas"] = 2.0 * n ** 3 / (best(lambda: torch.matmul(a32, b32), 3) * 1e-3) / 1e12
    src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    out["hbm_copy_gbs"] = 2.0 * src.numel() / (best(lambda: dst.copy_(src)) * 1e-3) / 1e9
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.s2d_debug_ffma.restype = ctypes.c_longlong
    lib.s2d_debug_ffma.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    sink = torch.zeros(1, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    flops = lib.s2d_debug_ffma(1, sink.data_ptr(), st)
    iters = 4096
    ms = best(lambda: lib.s2d_debug_ffma(iters, sink.data_ptr(), st))
    out["fp32_ffma_tflops"] = flops * iters / (ms * 1e-3) / 1e12
    out["nominal_fp32_ffma_tflops"] = 148 * 128 * 2 * 1.965e9 / 1e12
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
