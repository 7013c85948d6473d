#!/usr/bin/env python
"""Development aid: per-step beam candidates of the product path next to the oracle's (tests/test_gpu_beam.py case nb=4, eos x4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import synth  # noqa: E402
from radialog_b200.llm import LlamaForCausalLM  # noqa: E402
from oracle import radialog_oracle as O  # noqa: E402

nb, boost = 4, 4.0
dev = torch.device("cuda:0")
cfg = synth.tiny_llama_cfg()
sd = synth.make_llama_weights(cfg, seed=3, dtype=torch.float32)
sd = {k: v.to(torch.float16).float() for k, v in sd.items()}
sd["lm_head.weight"][cfg.eos_token_id] *= boost
model = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=torch.float16, device=dev)
orc = O.LlamaOracle(cfg, sd, torch.float16)
prompts = synth.make_prompts(3, seed=50 + nb, ragged=True)
prompts = torch.where(prompts == synth.IMG_TOKEN_ID, torch.full_like(prompts, 77), prompts)
log = {"o": [], "p": []}
_topk = torch.sort


def spy(which):
    def f(x, *a, **kw):
        r = _topk(x, *a, **kw)
        if x.dim() == 2 and x.shape[0] == 3:
            log[which].append((r.values[:, :2 * nb].detach().cpu().clone(), r.indices[:, :2 * nb].detach().cpu().clone()))
        return r
    return f


torch.sort = spy("o")
o_seq, o_sc = O.llama_beam_search(orc, prompts, None, 14, nb)
torch.sort = spy("p")
out = model.generate(prompts.to(dev), max_new_tokens=14, num_beams=nb, return_dict_in_generate=True)
torch.sort = _topk
print("oracle scores", o_sc.tolist(), "product", out.sequences_scores.tolist())
V = cfg.vocab_size
for s in range(min(len(log["o"]), len(log["p"]))):
    ov, oi = log["o"][s]
    pv, pi = log["p"][s]
    for b in range(3):
        print(f"step {s} row {b}: oracle", [(int(i) // V, int(i) % V, round(float(v), 3)) for v, i in zip(ov[b], oi[b])])
        print(f"step {s} row {b}: produc", [(int(i) // V, int(i) % V, round(float(v), 3)) for v, i in zip(pv[b], pi[b])])
print(o_seq[:, prompts.shape[1]:])
print(out.sequences.cpu()[:, prompts.shape[1]:])
