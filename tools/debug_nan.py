#!/usr/bin/env python
"""Development aid: which rd_linear algo leaves NaN/unwritten outputs for the SwiGLU decode shape."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
dtype = torch.float16
ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)


def run(x, w, M, N, K, algo, act, splits=0):
    out = torch.full((M, N), float("nan"), device=dev, dtype=dtype)
    e = _lib.Epilogue()
    e.act = act
    e.res_mode = 1
    lib.rd_linear_force_splits(splits)
    st = lib.rd_linear(x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, C.byref(e), 0, algo, ws.data_ptr(), ws.numel(),
                       torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rd_linear")
    torch.cuda.synchronize()
    return out


for (N, K, act) in [(11008, 4096, 3), (5504, 4096, 3), (8192, 4096, 3), (11008, 4096, 0), (11008, 1024, 3)]:
    rows = 2 * N if act else N
    g = torch.Generator().manual_seed(N)
    x = (torch.randn(32, K, generator=g) * 0.5).to(dtype).to(dev)
    w = (torch.randn(rows, K, generator=g) * 0.05).to(dtype).to(dev)
    xf, wf = x.float(), w.float()
    if act:
        gg = (xf @ wf[:N].t()).to(dtype)
        uu = (xf @ wf[N:].t()).to(dtype)
        ref = torch.nn.functional.silu(gg.float()).to(dtype) * uu
    else:
        ref = (xf @ wf.t()).to(dtype)
    for name, algo, sp in [("tc s=heur", 2, 0), ("tc s=1", 2, 1), ("tc s=2", 2, 2), ("simt", 3, 0)]:
        o = run(x, w, 32, N, K, algo, act, sp)
        nan = torch.isnan(o)
        cols = nan.any(0).nonzero().flatten()
        bad = (~nan) & ((o.float() - ref.float()).abs() > 0.02 * ref.float().abs().max())
        print(f"N={N} K={K} act={act} {name:10s}: nan elems {int(nan.sum())} cols [{cols.min().item() if len(cols) else '-'}, "
              f"{cols.max().item() if len(cols) else '-'}] ncols {len(cols)}; wrong (non-nan) elems {int(bad.sum())}", flush=True)
lib.rd_linear_force_splits(0)
