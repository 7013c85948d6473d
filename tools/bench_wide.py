#!/usr/bin/env python
"""Wide-token GEMM timing (development tool; run under gpurun): the Vicuna-7B prefill shapes (2048 tokens) and a few ResNet-50
convolution shapes through rd_linear, persistent kernel (linear_wide.cu) on / off and per token-tile width.  Prints TFLOP/s
against the measured bf16 peak.  usage: bench_wide.py [--nt 0,256,240] [--iters 20]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radialog_b200 import _lib  # noqa: E402

SHAPES = [
    # name, M, N, K, act, residual
    ("qkv", 2048, 12416, 4096, 0, False),
    ("o", 2048, 4096, 4096, 0, True),
    ("gate_up", 2048, 11008, 4096, _lib.ACT_SWIGLU, False),
    ("down", 2048, 4096, 11008, 0, True),
    ("l1_conv3", 100352, 256, 64, _lib.ACT_RELU, True),
    ("l1_conv1", 100352, 64, 256, _lib.ACT_RELU, False),
    ("l2_conv2", 25088, 128, 1152, _lib.ACT_RELU, False),
    ("l3_conv3", 6272, 1024, 256, _lib.ACT_RELU, True),
    ("l4_conv2", 1568, 512, 4608, _lib.ACT_RELU, False),
    ("qf_self_qkv", 1024, 2304, 768, _lib.ACT_RELU, False),
    ("qf_dense", 1024, 768, 768, _lib.ACT_RELU, True),
    ("qf_fc1", 1024, 3072, 768, _lib.ACT_RELU, False),
    ("qf_fc2", 1024, 768, 3072, _lib.ACT_RELU, True),
    ("qf_cross_kv", 6272, 9216, 1408, _lib.ACT_RELU, False),
    ("proj", 6272, 1408, 1408, _lib.ACT_RELU, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", default="0")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shapes", default="")
    ap.add_argument("--stages", default="0", help="comma list of pipeline depths to try per nt (0 = by K)")
    ap.add_argument("--sustained", type=int, default=0, help="N > 0: time N back-to-back launches with no L2 flush in between (power-limited "
                    "clocks, as inside a prefill) and print torch.matmul (cuBLAS) on the same shape beside it")
    ap.add_argument("--min-tiles", type=int, default=1)
    args = ap.parse_args()
    lib = _lib.load()
    lib.rd_linear_wide_min_tiles(args.min_tiles)
    dev = torch.device("cuda:0")
    dtype = torch.bfloat16
    peak = 1404.0
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, M, N, K, act, res in SHAPES:
        if args.shapes and name not in args.shapes.split(","):
            continue
        rows = 2 * N if act == _lib.ACT_SWIGLU else N
        w = (torch.randn(rows, K, device=dev) * 0.02).to(dtype)
        x = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
        residual = (torch.randn(M, N, device=dev) * 0.5).to(dtype) if res else None
        bias = torch.randn(N, device=dev) * 0.1 if act == _lib.ACT_RELU else None
        out = torch.empty(M, N, device=dev, dtype=dtype)
        ws = torch.zeros(int(lib.rd_linear_workspace_bytes(M, N, K)) + 256, dtype=torch.uint8, device=dev)
        e = _lib.Epilogue()
        e.bias_dev = _lib.ptr(bias)
        e.residual_dev = _lib.ptr(residual)
        e.ld_res = N
        e.res_mode = 2 if act == _lib.ACT_RELU else 1
        e.act = act
        flops = 2.0 * M * rows * K
        line = [f"{name:9s} M={M:6d} N={N:5d} K={K:5d}"]
        configs = [("old", 0, 0, 0, 0)] + [(f"{'pair' if pr else 'single'} nt={nt} s={sg}", 1, int(nt), int(sg), pr) for pr in (0, 1)
                                          for nt in args.nt.split(",") for sg in args.stages.split(",")]
        for label, on, nt, sg, pr in configs:
            lib.rd_linear_wide_persistent(on)
            lib.rd_linear_wide_pair(pr)
            lib.rd_linear_wide_force_nt(nt)
            lib.rd_linear_wide_force_stages(sg)
            st = _lib.current_stream()

            def run():
                _lib.check(lib.rd_linear(_lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(out), N, M, N, K, C.byref(e), _lib.dtype_code(dtype), 2,
                                         _lib.ptr(ws), ws.numel(), st), "rd_linear")
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts = []
            if args.sustained > 0:
                for _ in range(args.sustained):      # bring the chip to its sustained state first
                    run()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(args.sustained):
                    run()
                b.record()
                torch.cuda.synchronize()
                ts = [a.elapsed_time(b) * 1e3 / args.sustained]
            for _ in range(args.iters if args.sustained == 0 else 0):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            us = ts[len(ts) // 2]
            line.append(f"{label}: {us:7.1f} us ({flops / us / 1e6 / peak:.2f})")
        if args.sustained > 0:
            wt = w.t().contiguous()
            for _ in range(args.sustained):
                torch.matmul(x, wt)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.sustained):
                torch.matmul(x, wt)
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / args.sustained
            line.append(f"cuBLAS (no epilogue): {us:7.1f} us ({flops / us / 1e6 / peak:.2f})")
        lib.rd_linear_wide_force_stages(0)
        lib.rd_linear_wide_pair(1)
        lib.rd_linear_wide_persistent(1)
        lib.rd_linear_wide_force_nt(0)
        print("  |  ".join(line), flush=True)


if __name__ == "__main__":
    main()
