#!/usr/bin/env python
"""ncu target: every kernel class of the hot path at its real shape, few launches (profiles/ncu_r2_*.md).

    python tools/ncu_step_target.py llm [B] [dtype]     # full-width Vicuna-7B layers (2 of them): one prefill + two decode steps
    python tools/ncu_step_target.py vision [B]          # full ResNet-50 + Q-Former forward of B images

Run under `ncu --set full --clock-control none --import-source on`; numbers printed by a run under ncu are not bench values.
Kernel shapes do not depend on the layer count, so the LLM target keeps 2 layers to stay within a short ncu session."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "llm"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
lib = _lib.load()
lib.rd_set_pdl(1)
if what == "llm":
    from radialog_b200.llm import LlamaForCausalLM
    dtype = getattr(torch, sys.argv[3]) if len(sys.argv) > 3 else torch.bfloat16
    cfg = synth.LlamaCfg(num_hidden_layers=2)
    sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
    llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
    llm.use_cuda_graph = False
    prompts = synth.make_prompts(B, seed=4321).to(dev)
    img = torch.randn(B, 32, 768, device=dev) * 0.5
    # context ~128 as in the middle of a 128-token report: prompt 64 + 64 generated
    out = llm.generate(prompts, img_embeds=img, max_new_tokens=3, suppress_eos=True)
    conv = torch.cat([out, torch.randint(3, 32000, (B, 61), device=dev)], -1)
    llm.generate(conv, img_embeds=img, max_new_tokens=3, suppress_eos=True)
    torch.cuda.synchronize()
else:
    from radialog_b200.vision import Blip2Qformer
    vcfg = synth.VisionCfg()
    vis = Blip2Qformer.from_state_dict(vcfg, synth.make_vision_weights(vcfg, seed=0), torch_dtype=torch.float16, device=dev, max_batch=B)
    imgs = synth.make_images(B, seed=1234).to(dev)
    vis.forward_image(imgs)
    torch.cuda.synchronize()
print("done")
