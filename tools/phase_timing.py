#!/usr/bin/env python
"""Vision and prefill phase timing at the bench shapes, with the per-kernel-class split of the prefill (development tool)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib, synth  # noqa: E402
from radialog_b200.llm import LlamaForCausalLM  # noqa: E402
from radialog_b200.vision import Blip2Qformer  # noqa: E402

dev = torch.device("cuda:0")
dtype = torch.bfloat16
lib = _lib.load()
lib.rd_set_pdl(int(os.environ.get("PDL", "1")))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if os.environ.get("SKIP_VISION") is None:
    vcfg = synth.VisionCfg()
    vis = Blip2Qformer.from_state_dict(vcfg, synth.make_vision_weights(vcfg, seed=0), torch_dtype=dtype, device=dev, max_batch=B)
    imgs = synth.make_images(B, seed=1234).to(dev)
    ms = timed(lambda: vis.forward_image(imgs))
    flops = B * (33.96e9 + 11.1e9)
    print(f"vision B={B}: {ms:.2f} ms  ({flops / ms / 1e9:.0f} TFLOP/s of algorithmic work)", flush=True)

cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
prompts = synth.make_prompts(B, seed=4321).to(dev)
img = torch.randn(B, 32, 768, device=dev) * 0.5
llm.reserve(B, 256)
ms = timed(lambda: llm.generate(prompts, img_embeds=img, max_new_tokens=1, suppress_eos=True))
flops = 2 * B * 64 * 32 * 202_375_168
print(f"prefill B={B} T=64: {ms:.2f} ms  ({flops / ms / 1e9:.0f} TFLOP/s)", flush=True)
_lib.check(lib.rd_llm_profile(llm._h, 1), "profile")
llm.generate(prompts, img_embeds=img, max_new_tokens=1, suppress_eos=True)
n = len(_lib.PROFILE_CLASSES)
msa = (C.c_float * n)()
cnt = (C.c_int * n)()
_lib.check(lib.rd_llm_profile_read(llm._h, msa, cnt, n), "read")
lib.rd_llm_profile(llm._h, 0)
print("prefill classes:", {k: (round(float(msa[i]), 2), int(cnt[i])) for i, k in enumerate(_lib.PROFILE_CLASSES) if cnt[i]})
