#!/usr/bin/env python
"""Per-kernel-class time of eager decode steps (CUDA events around each launch), with the o_proj / down_proj partials finished
by the norm kernel (default) and by the GEMMs' own cluster reduction (development tool).  python tools/class_timing.py [B]"""
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
lib.rd_set_pdl(1)
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
modes = [("od_partials", lambda: llm.set_od_partials(True)),
         ("cluster_reduce", lambda: llm.set_od_partials(False))]
for sk, fn in modes:
    fn()
    prof = llm.profile_decode_steps(B, 64, steps=8)
    print(f"B={B} {sk}: " + "  ".join(f"{k} {v['ms'] / max(1, v['launches']) * 1e3:.1f}us x{v['launches'] // 8}" for k, v in prof.items() if v["launches"]))
