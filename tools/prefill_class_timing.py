#!/usr/bin/env python
"""Per-kernel-class device time of one B x T prefill (CUDA events around every launch; development tool)."""
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib.rd_set_pdl(1)
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
ids = synth.make_prompts(B, seed=1).to(dev).contiguous()
llm.reserve(B, 64 + 130)
st = _lib.current_stream()
for rep in range(2):
    _lib.check(lib.rd_llm_profile(llm._h, 1 if rep == 1 else 0), "profile")
    _lib.check(lib.rd_llm_prefill(llm._h, _lib.ptr(ids), None, B, 64, None, 1, st), "prefill")
    torch.cuda.synchronize()
n = len(_lib.PROFILE_CLASSES)
ms = (C.c_float * n)()
cnt = (C.c_int * n)()
_lib.check(lib.rd_llm_profile_read(llm._h, ms, cnt, n), "read")
print("prefill B=%d T=64: " % B + "  ".join(f"{k} {ms[i]:.2f}ms/{cnt[i]}" for i, k in enumerate(_lib.PROFILE_CLASSES) if cnt[i]) + f"  total {sum(ms):.1f} ms")
