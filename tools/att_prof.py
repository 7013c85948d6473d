#!/usr/bin/env python
"""Per-phase clock64 profile of attention_decode_kernel inside real decode steps (build with RD_EXTRA_NVCC_FLAGS=-DRD_ATT_PROF)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radialog_b200 import _lib, synth
from radialog_b200.llm import LlamaForCausalLM
dev = torch.device("cuda:0")
dtype = torch.bfloat16
cfg = synth.LlamaCfg(num_hidden_layers=4)
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
B = 32
prompts = synth.make_prompts(B, seed=4321).to(dev)
img = torch.randn(B, 32, 768, device=dev) * 0.5
out = llm.generate(prompts, img_embeds=img, max_new_tokens=100, suppress_eos=True)
torch.cuda.synchronize()
