"""Throughput of this repo's LLM path against the reference's GPU PyTorch path on the same B200 (north_star: ">= 10x the
reference GPU PyTorch path's reports/sec at batch 32").

The reference stack (torch 1.13 / transformers 4.28.1 / peft) is not installable here, so its GPU path is the oracle's
restatement of it (oracle/radialog_oracle.py: eager PyTorch ops, three separate q/k/v Linears, unmerged LoRA, torch.cat KV
cache, all-position lm_head, per-token Python loop - SURVEY.md 8d item 2) run on cuda with the model dtype's cuBLAS GEMMs,
on the same seeded full-size weights.  Only the LLM part (prefill + greedy decode) is timed; it is >= 95 % of a report."""
import pytest
import torch
import torch.nn.functional as F

from radialog_b200 import synth
from radialog_b200.llm import LlamaForCausalLM
from oracle import radialog_oracle as O

pytestmark = pytest.mark.gpu


def test_batch32_speedup_over_reference_gpu_pytorch_path(cuda_dev, monkeypatch, capsys):
    dtype = torch.bfloat16
    cfg = synth.LlamaCfg()
    B, T, NEW = 32, 64, 32
    sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device=str(cuda_dev))
    model = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=cuda_dev)
    prompts = synth.make_prompts(B, seed=4321).to(cuda_dev)
    g = torch.Generator().manual_seed(7)
    img = (torch.randn(B, 32, cfg.qformer_hidden, generator=g) * 0.5).to(cuda_dev)

    def timed(fn):
        fn()                                     # warm-up (graph capture / cuBLAS heuristics)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    ms_own, ids_own = timed(lambda: model.generate(prompts, img_embeds=img, max_new_tokens=NEW, suppress_eos=True))

    # reference path: the oracle's ops on cuda, GEMMs through cuBLAS in the model dtype (what nn.Linear does on the GPU)
    monkeypatch.setattr(O, "_mm", lambda x, w, dt: F.linear(x, w))
    orc = O.LlamaOracle(cfg, sd, dtype)
    orc.cos, orc.sin = orc.cos.to(cuda_dev), orc.sin.to(cuda_dev)

    def ref():
        with torch.device(cuda_dev), torch.no_grad():
            return orc.generate(prompts, img, NEW, suppress_eos=True)

    ms_ref, ids_ref = timed(ref)
    ratio = ms_ref / ms_own
    same = (ids_own[:, :T + NEW].cpu() == ids_ref.cpu()).float().mean().item()
    with capsys.disabled():
        print(f"\n[reference GPU PyTorch path] B={B} T={T} {NEW} new tokens: reference {ms_ref:.0f} ms, this repo {ms_own:.0f} ms "
              f"-> {ratio:.1f}x ({B * 1e3 / ms_ref:.1f} vs {B * 1e3 / ms_own:.1f} reports/s at this length), token agreement {same:.3f}")
    assert ratio >= 3.0, f"only {ratio:.2f}x over the eager PyTorch path"
