#!/usr/bin/env python
"""Decode-step time (CUDA graph + PDL, the bench configuration) of the full Vicuna-7B-sized engine under engine switches
(development tool): TMEM staging, split-K partial hand-offs, L2 weight prefetch.  python tools/decode_sweep.py [B] [NEW] [out.json]"""
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
MBY = 1 << 20


def apply(ts=0, od=1, qp=1, pf=(0, 0, 0)):
    lib.rd_linear_tmem_staging(ts)
    llm.set_od_partials(bool(od))
    llm.set_qkv_partials(bool(qp))
    llm._graphs = {}
    _lib.check(lib.rd_llm_set_l2_prefetch(llm._h, pf[0], pf[1], pf[2]), "l2 prefetch")


variants = [("default", dict()), ("tmem staging", dict(ts=1)), ("od_partials off", dict(od=0)), ("qkv_partials off", dict(qp=0)),
            ("l2 prefetch o=32MB gate_up=32MB (from the attention kernel)", dict(pf=(0, 32 * MBY, 32 * MBY))),
            ("l2 prefetch qkv=32MB (from the norm kernel: inactive with od_partials)", dict(pf=(32 * MBY, 0, 0))),
            ("default (again)", dict())]
res = []
llm.generate(prompts, img_embeds=img, max_new_tokens=4, suppress_eos=True)      # creates the engine handle
for name, kw in variants:
    apply(**kw)
    llm.generate(prompts, img_embeds=img, max_new_tokens=8, suppress_eos=True)
    best = 1e9
    for _ in range(2):
        llm.generate(prompts, img_embeds=img, max_new_tokens=NEW, suppress_eos=True)
        best = min(best, llm.last_stats["decode_ms"] / (NEW - 1))
    res.append({"variant": name, "B": B, "ms_per_step": best})
    print(f"B={B} {name:70s} {best:.3f} ms/step", flush=True)
apply()
if len(sys.argv) > 3:
    json.dump(res, open(sys.argv[3], "w"), indent=1)
