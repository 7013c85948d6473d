#!/usr/bin/env python
"""Timeline of the GEMM launches inside one eager decode step of the Vicuna-7B-sized engine (development tool):
per launch first CTA entry, last CTA end, and the gap to the previous GEMM (= the non-GEMM kernels + launch latency)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib, synth  # noqa: E402
from radialog_b200.llm import LlamaForCausalLM  # noqa: E402

dev = torch.device("cuda:0")
dtype = torch.bfloat16
lib = _lib.load()
lib.rd_linear_set_trace_strided.argtypes = [C.c_void_p, C.c_longlong]
lib.rd_linear_trace_launches.restype = C.c_longlong
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib.rd_set_pdl(int(os.environ.get("PDL", "1")))
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
llm.use_cuda_graph = False
if os.environ.get("ALGO"):
    llm.set_algo(int(os.environ["ALGO"]))
prompts = synth.make_prompts(B, seed=4321).to(dev)
img = torch.randn(B, 32, 768, device=dev) * 0.5
llm.generate(prompts, img_embeds=img, max_new_tokens=6, suppress_eos=True)
torch.cuda.synchronize()
STRIDE = 512 * 16
trace = torch.zeros(400 * STRIDE, dtype=torch.int64, device=dev)
lib.rd_linear_set_trace_strided(trace.data_ptr(), STRIDE)
st = _lib.current_stream()
for _ in range(2):
    _lib.check(lib.rd_llm_decode_step(llm._h, st), "decode_step")
torch.cuda.synchronize()
n = lib.rd_linear_trace_launches()
lib.rd_linear_set_trace(None)
t = trace.view(-1, 512, 16)[:n].cpu().double()
per = n // 2
rows = []
for i in range(per, n):          # second step
    x = t[i]
    x = x[x[:, 0] > 0]
    rows.append((x[:, 0].min().item(), x[:, 3].median().item(), x[:, 4].max().item(), x[:, 7].max().item(), x.shape[0]))
t0 = rows[0][0]
names = ["qkv", "o", "gate_up", "down"]
print(f"{per} GEMM launches per step; step GEMM span {(rows[-1][3] - t0) / 1e3:.1f} us")
acc = {}
prev_end = None
for i, (s0, fd, lm, e, nc) in enumerate(rows):
    nm = names[i % 4] if i < per - 1 else "lm_head"
    gap = (s0 - prev_end) / 1e3 if prev_end else 0.0
    d = acc.setdefault(nm, [0, 0.0, 0.0, 0.0, 0.0])
    d[0] += 1; d[1] += (e - s0) / 1e3; d[2] += gap; d[3] += (fd - s0) / 1e3; d[4] += (e - lm) / 1e3
    prev_end = e
    if 4 <= i < 8:
        print(f"  layer1 {nm:8s} ctas {nc:4d} start {(s0 - t0) / 1e3:8.1f} first_data +{(fd - s0) / 1e3:5.1f} last_mma +{(lm - s0) / 1e3:5.1f} end +{(e - s0) / 1e3:5.1f}  gap_before {gap:5.1f}")
for nm, (c, dur, gap, fd, tail) in acc.items():
    print(f"{nm:8s} x{c:3d}: kernel span {dur / c:6.1f} us  (first data +{fd / c:4.1f}, tail after last MMA {tail / c:4.1f})   gap before {gap / c:6.1f} us")
