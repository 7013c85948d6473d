#!/usr/bin/env python
"""Implicit-GEMM convolution vs im2col + GEMM on every conv geometry of the ResNet-50 trunk (development tool)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from radialog_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
dtype = torch.float16
geoms = [(112, 64, 64, 3, 1), (112, 128, 128, 3, 2), (56, 128, 128, 3, 1), (56, 256, 256, 3, 2), (28, 256, 256, 3, 1), (28, 512, 512, 3, 2),
         (14, 512, 512, 3, 1), (112, 256, 512, 1, 2), (56, 512, 1024, 1, 2), (28, 1024, 2048, 1, 2)]
for B in (1, 2, 32):
    for (hw, cin, cout, ks, stride) in geoms:
        pad = 1 if ks == 3 else 0
        g = torch.Generator().manual_seed(hw + cin)
        x = (torch.randn(B, hw, hw, cin, generator=g) * 0.5).to(dtype).to(dev)
        w = (torch.randn(cout, ks * ks * cin, generator=g) * 0.05).to(dtype).to(dev)
        bias = torch.randn(cout, generator=g).to(dev) * 0.1
        oh = (hw + 2 * pad - ks) // stride + 1
        M, K = B * oh * oh, ks * ks * cin
        e = _lib.Epilogue()
        e.bias_dev = _lib.ptr(bias)
        e.act = _lib.ACT_RELU
        e.res_mode = 2
        out_i = torch.full((M, cout), float("nan"), device=dev, dtype=dtype)
        r = lib.rd_conv_nhwc_implicit(_lib.ptr(x), _lib.ptr(w), _lib.ptr(out_i), cout, B, hw, hw, cin, cout, ks, stride, pad, C.byref(e),
                                      _lib.dtype_code(dtype), _lib.current_stream())
        torch.cuda.synchronize()
        col = torch.empty(M, K, device=dev, dtype=dtype)
        _lib.check(lib.rd_im2col_nhwc(_lib.ptr(x), _lib.ptr(col), B, hw, hw, cin, ks, stride, pad, _lib.dtype_code(dtype), _lib.current_stream()), "im2col")
        out_e = torch.empty(M, cout, device=dev, dtype=dtype)
        _lib.check(lib.rd_linear(_lib.ptr(col), K, _lib.ptr(w), K, _lib.ptr(out_e), cout, M, cout, K, C.byref(e), _lib.dtype_code(dtype), 2, None, 0,
                                 _lib.current_stream()), "linear")
        torch.cuda.synchronize()
        d = (out_i.float() - out_e.float()).abs()
        bad = (d > 0.01) | ~torch.isfinite(out_i.float())
        rows = bad.any(dim=1).nonzero().flatten()
        print(f"B={B} hw={hw} cin={cin} cout={cout} ks={ks} s={stride}: r={r} M={M} max diff {d.max().item():.4g} bad rows {rows.numel()}"
              + (f" first {rows[:4].tolist()} last {rows[-4:].tolist()}" if rows.numel() else ""), flush=True)
