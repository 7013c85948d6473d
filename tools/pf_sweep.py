#!/usr/bin/env python
"""Decode-step time vs L2 weight-prefetch budgets (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radialog_b200 import _lib, synth
from radialog_b200.llm import LlamaForCausalLM
dev = torch.device("cuda:0"); dtype = torch.bfloat16
lib = _lib.load(); lib.rd_set_pdl(1)
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev); del sd
for B in (32, 1):
    prompts = synth.make_prompts(B, seed=4321).to(dev); img = torch.randn(B, 32, 768, device=dev) * 0.5
    llm.generate(prompts, img_embeds=img, max_new_tokens=4, suppress_eos=True)
    for (q, o, g) in [(0, 0, 0), (64, 34, 48), (0, 34, 48), (32, 34, 32), (48, 34, 90), (0, 34, 90), (64, 34, 0), (0, 0, 0)]:
        _lib.check(lib.rd_llm_set_l2_prefetch(llm._h, q << 20, o << 20, g << 20), "pf")
        llm._graphs = {}
        llm.generate(prompts, img_embeds=img, max_new_tokens=8, suppress_eos=True)
        llm.generate(prompts, img_embeds=img, max_new_tokens=64, suppress_eos=True)
        print(f"B={B:2d} prefetch qkv/o/gate_up = {q}/{o}/{g} MB: decode {llm.last_stats['decode_ms'] / 63:.3f} ms/step", flush=True)
