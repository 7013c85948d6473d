#!/usr/bin/env python
"""Tiny ncu target: a few eager decode steps of the Vicuna-7B-sized engine through the persistent kernel
(development tool).  python tools/ncu_mega_target.py [B] [steps]"""
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
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib.rd_set_pdl(1)
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
llm.use_cuda_graph = False
llm.set_mega(True)
prompts = synth.make_prompts(B, seed=4321).to(dev)
img = torch.randn(B, 32, 768, device=dev) * 0.5
llm.reserve(B, 64 + steps + 8)
llm.generate(prompts, img_embeds=img, max_new_tokens=steps + 2, suppress_eos=True)
torch.cuda.synchronize()
