#!/usr/bin/env python
"""Decode-step time (CUDA graph + PDL, the bench configuration) of the full Vicuna-7B-sized engine under engine switches
(development tool): o_proj / down_proj partials finished by the norm launch on/off.  python tools/decode_sweep.py [B] [NEW]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib, synth  # noqa: E402
from radialog_b200.llm import LlamaForCausalLM  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NEW = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
dtype = torch.bfloat16
lib = _lib.load()
lib.rd_set_pdl(1)
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
torch.cuda.empty_cache()
prompts = synth.make_prompts(B, seed=4321).to(dev)
img = torch.randn(B, 32, 768, device=dev) * 0.5
MB = 1 << 20
variants = [("od_partials=1", 1), ("od_partials=0", 0), ("od_partials=1 (again)", 1), ("od_partials=0 (again)", 0)]
res = []
for name, ts in variants:
    llm.set_od_partials(bool(ts))
    llm.generate(prompts, img_embeds=img, max_new_tokens=8, suppress_eos=True)
    best = 1e9
    for _ in range(2):
        llm.generate(prompts, img_embeds=img, max_new_tokens=NEW, suppress_eos=True)
        best = min(best, llm.last_stats["decode_ms"] / (NEW - 1))
    res.append({"variant": name, "B": B, "ms_per_step": best})
    print(f"B={B} {name:40s} {best:.3f} ms/step", flush=True)
if len(sys.argv) > 3:
    json.dump(res, open(sys.argv[3], "w"), indent=1)
