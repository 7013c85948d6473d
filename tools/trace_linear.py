#!/usr/bin/env python
"""Per-CTA timeline of linear_tc_kernel on the decode shapes (development tool): where a CTA's life goes.
Stamps (globaltimer, ns): 0 entry, 1 prologue done, 2 producer past pdl_wait, 3 first stage landed, 4 last MMA issued,
5 accumulators complete (epilogue), 6 after cluster barrier 1, 7 end."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib  # noqa: E402

SH = {"qkv": (12288, 4096, 0), "o": (4096, 4096, 0), "gate_up": (11008, 4096, 3), "down": (4096, 11008, 0), "lm_head": (32001, 4096, 0)}
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib = _lib.load()
lib.rd_linear_set_trace.argtypes = [C.c_void_p]
lib.rd_set_pdl(int(os.environ.get("PDL", "0")))
dev = torch.device("cuda:0")
ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
trace = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, (N, K, act) in SH.items():
    rows = 2 * N if act else N
    w = (torch.randn(rows, K, device=dev) * 0.02).half()
    x = (torch.randn(M, K, device=dev) * 0.5).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    e = _lib.Epilogue()
    e.act = act
    e.res_mode = 1

    def launch():
        _lib.check(lib.rd_linear(x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, C.byref(e), 0, 2, ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream().cuda_stream), "rd_linear")

    launch()
    flush.zero_()
    torch.cuda.synchronize()
    trace.zero_()
    lib.rd_linear_set_trace(trace.data_ptr())
    launch()
    torch.cuda.synchronize()
    lib.rd_linear_set_trace(None)
    t = trace.view(-1, 16).cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    names = ["entry", "prologue", "pdl_wait", "first_data", "last_mma", "acc_done", "cl_bar1", "end", "ldtm0", "st16a", "st16b", "epi_done"]
    print(f"{name} M={M}: {t.shape[0]} CTAs, kernel span {rel[:, 7].max():.1f} us (ideal {rows * K * 2 / 6551.7e3:.1f} us)")
    for i, nm in enumerate(names):
        col = rel[:, i]
        col = col[t[:, i] > 0]
        if len(col):
            print(f"   {nm:10s} min {col.min():7.2f}  median {col.median():7.2f}  max {col.max():7.2f} us")
    pairs = [(1, 3), (3, 4), (4, 5), (5, 6), (6, 7), (5, 7), (5, 8), (8, 9), (9, 10), (10, 11), (11, 7)]
    for a, b in pairs:
        ok = (t[:, a] > 0) & (t[:, b] > 0)
        if ok.any():
            d = (t[ok, b] - t[ok, a]) / 1e3
            print(f"   delta {names[a]:>10s} -> {names[b]:10s} min {d.min():6.2f} median {d.median():6.2f} max {d.max():6.2f} us")
