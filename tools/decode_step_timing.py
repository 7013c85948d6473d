#!/usr/bin/env python
"""Decode-step time of the full Vicuna-7B-sized engine for {graph, eager} x {PDL off, on} (development tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib, synth  # noqa: E402
from radialog_b200.llm import LlamaForCausalLM  # noqa: E402

dev = torch.device("cuda:0")
dtype = torch.bfloat16
lib = _lib.load()
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
torch.cuda.empty_cache()
NEW = int(os.environ.get("NEW", "64"))
for B in [int(b) for b in os.environ.get("BS", "32,1").split(",")]:
    prompts = synth.make_prompts(B, seed=4321).to(dev)
    img = torch.randn(B, 32, 768, device=dev) * 0.5
    for algo in ([0, 1, 2] if B <= 4 else [0]):
      llm.set_algo(algo)
      for graph, pdl in ((True, 1), (False, 1), (True, 0)):
        if True:
            lib.rd_set_pdl(pdl)
            llm.use_cuda_graph = graph
            llm._graphs = {}
            llm.generate(prompts, img_embeds=img, max_new_tokens=8, suppress_eos=True)
            torch.cuda.synchronize()
            llm.generate(prompts, img_embeds=img, max_new_tokens=NEW, suppress_eos=True)
            s = llm.last_stats
            print(f"B={B:3d} algo={algo} graph={int(graph)} pdl={pdl}: prefill {s['prefill_ms']:.1f} ms, decode {s['decode_ms'] / (NEW - 1):.3f} ms/step", flush=True)
lib.rd_set_pdl(0)
