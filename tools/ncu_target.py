#!/usr/bin/env python
"""Tiny ncu target: a few launches of rd_linear on chosen decode shapes (development tool).
python tools/ncu_target.py qkv,gate_up 32 [pdl]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib  # noqa: E402

SH = {"qkv": (12288, 4096, 0), "o": (4096, 4096, 0), "gate_up": (11008, 4096, 3), "down": (4096, 11008, 0), "lm_head": (32001, 4096, 0)}
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["qkv", "gate_up"]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lib = _lib.load()
lib.rd_set_pdl(1 if len(sys.argv) > 3 else 0)
dev = torch.device("cuda:0")
ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
for name in names:
    N, K, act = SH[name]
    rows = 2 * N if act else N
    Ws = [(torch.randn(rows, K, device=dev) * 0.02).half() for _ in range(3)]
    x = (torch.randn(M, K, device=dev) * 0.5).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    e = _lib.Epilogue()
    e.act = act
    e.res_mode = 1
    for i in range(4):
        st = lib.rd_linear(x.data_ptr(), K, Ws[i % 3].data_ptr(), K, out.data_ptr(), N, M, N, K, C.byref(e), 0, 0, ws.data_ptr(), ws.numel(),
                           torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "rd_linear")
    torch.cuda.synchronize()
    del Ws
