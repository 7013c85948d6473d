#!/usr/bin/env python
"""Kernel-level roofline sweep of rd_linear on the Vicuna-7B decode shapes (development tool; run under gpurun).

For each (shape, M, algo, splits) it cycles through enough distinct weight buffers to exceed the 126 MB L2, times the
launches with CUDA events on the launching stream and prints achieved GB/s of algorithmic weight bytes (2*N*K) next to
the measured HBM peak.  python tools/bench_linear.py [--pdl 1] [--json gpurun_out/linear_sweep.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib  # noqa: E402

SHAPES = [("qkv", 12288, 4096, 0), ("o", 4096, 4096, 0), ("gate_up", 11008, 4096, 3), ("down", 4096, 11008, 0), ("lm_head", 32001, 4096, 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pdl", type=int, default=0)
    ap.add_argument("--json", default="")
    ap.add_argument("--ms", default="1,4,8,32")
    ap.add_argument("--splits", default="0")
    ap.add_argument("--dtype", default="float16")
    ap.add_argument("--splitk-mode", type=int, default=0, help="0 cluster/DSMEM reduction, 1 global workspace")
    ap.add_argument("--nbuf", type=int, default=0, help="weight buffers cycled through (0 = enough to exceed L2; 1 = L2-resident weights)")
    ap.add_argument("--shapes", default="", help="comma list of shape names (default all)")
    args = ap.parse_args()
    lib = _lib.load()
    lib.rd_set_pdl(args.pdl)
    lib.rd_linear_splitk_mode(args.splitk_mode)
    dev = torch.device("cuda:0")
    dtype = getattr(torch, args.dtype)
    peak = 6551.7
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    ws = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
    results = []
    for name, N, K, act in SHAPES:
        if args.shapes and name not in args.shapes.split(","):
            continue
        rows = 2 * N if act == 3 else N
        nbuf = args.nbuf if args.nbuf > 0 else max(2, int(400e6 // (rows * K * 2)) + 1)
        Ws = [(torch.randn(rows, K, device=dev) * 0.02).to(dtype) for _ in range(nbuf)]
        for M in [int(m) for m in args.ms.split(",")]:
            x = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
            out = torch.empty(M, N, device=dev, dtype=dtype)
            e = _lib.Epilogue()
            e.act = act
            e.res_mode = 1
            algos = [("gemv", 1)] if M <= 4 else []
            algos.append(("tc", 2))
            for aname, algo in algos:
                for sp in [int(s) for s in args.splits.split(",")]:
                    if algo == 1 and sp != 0:
                        continue
                    lib.rd_linear_force_splits(sp)

                    def launch(i):
                        w = Ws[i % nbuf]
                        st = lib.rd_linear(x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, C.byref(e), _lib.dtype_code(dtype),
                                           algo, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
                        _lib.check(st, "rd_linear")

                    for i in range(5):
                        launch(i)
                    torch.cuda.synchronize()
                    iters = 40
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(iters):
                        launch(i)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / iters
                    gbs = rows * K * 2 / (us * 1e-6) / 1e9
                    results.append(dict(shape=name, N=N, K=K, M=M, algo=aname, splits=sp, us=us, gbs=gbs, frac=gbs / peak))
                    print(f"{name:8s} M={M:3d} {aname:4s} splits={sp:2d}  {us:8.1f} us  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of measured peak", flush=True)
        del Ws
        torch.cuda.empty_cache()
    lib.rd_linear_force_splits(0)
    if args.json:
        json.dump(results, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
