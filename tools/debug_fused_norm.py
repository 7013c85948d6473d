#!/usr/bin/env python
"""Per-step logit difference between the fused-norm and the separate-norm decode paths (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib, synth  # noqa: E402
from radialog_b200.llm import LlamaForCausalLM  # noqa: E402

dev = torch.device("cuda:0")
dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = synth.tiny_llama_cfg(num_hidden_layers=3)
sd = {k: v.to(torch.float16).float() for k, v in synth.make_llama_weights(cfg, seed=0, dtype=torch.float32).items()}
m = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
prompts = synth.make_prompts(B, seed=299 + B, ragged=True).to(dev)
g = torch.Generator().manual_seed(B + 2)
img = (torch.randn(B, 32, cfg.qformer_hidden, generator=g) * 0.5).to(dev)
for graph in (False, True):
    m.use_cuda_graph = graph
    outs = {}
    for fused in (False, True):
        m.set_fused_norm(fused)
        outs[fused] = m.generate(prompts, img_embeds=img, max_new_tokens=9, suppress_eos=True, return_dict_in_generate=True, output_scores=True)
    a, b = outs[False], outs[True]
    errs = [(x.float() - y.float()).abs().max().item() for x, y in zip(a.scores, b.scores)]
    same = [(x.argmax(-1) == y.argmax(-1)).float().mean().item() for x, y in zip(a.scores, b.scores)]
    print(f"graph={int(graph)} max|dlogit| per step: " + " ".join(f"{e:.4f}" for e in errs))
    print(f"          argmax agreement per step: " + " ".join(f"{e:.2f}" for e in same))
