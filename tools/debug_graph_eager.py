import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from radialog_b200 import synth
from test_gpu_llm import build, img_tokens
dev = torch.device("cuda:0")
dtype = torch.float16
cfg = synth.tiny_llama_cfg(num_hidden_layers=3)
model, orc, _ = build(cfg, dtype, dev)
B = 32
prompts = synth.make_prompts(B, seed=777, ragged=True)
img = img_tokens(B, cfg, seed=5)
outs = []
for mode in (False, False, True, True, False):
    model.use_cuda_graph = mode
    outs.append(model.generate(prompts.to(dev), img_embeds=img.to(dev), max_new_tokens=10, suppress_eos=True).cpu())
T = prompts.shape[1]
for i in range(1, 5):
    d = (outs[0] != outs[i])
    print("run", i, "graph" if i in (2, 3) else "eager", "differs from run 0 at", d.nonzero()[:6].tolist(), "n", int(d.sum()))
o_ids, o_scores = orc.generate(prompts, img, 10, suppress_eos=True, return_scores=True)
for i in range(5):
    print("run", i, "vs oracle mismatches", int((outs[i] != o_ids).sum()))
