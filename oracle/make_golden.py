"""ORACLE tooling — container-only.  Generates ``tests/golden/*.npz`` by running the REFERENCE's own modules
(imported from /root/reference through ``ref_import``) on seeded synthetic weights/inputs, and reports how far
``radialog_oracle`` is from them.  Run:  ``python -m oracle.make_golden``  (from the repo root).

Fixtures hold only seeds/configs + the reference's outputs; weights are regenerated from the seed by
``radialog_b200.synth`` wherever the fixtures are consumed.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import radialog_oracle as O      # noqa: E402
from oracle import ref_import as R           # noqa: E402
from radialog_b200 import synth              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


def _np(t):
    t = t.detach()
    if t.dtype == torch.bfloat16:
        return t.float().numpy()
    return t.numpy()


def golden_llm(tag: str, cfg, dtype_name: str, B: int, new_tokens: int, ragged: bool, seed: int = 0):
    dtype = DT[dtype_name]
    torch.manual_seed(0)
    sd = synth.make_llama_weights(cfg, seed=seed, dtype=torch.float32)
    # values are representable in fp16; bf16/fp32 runs cast the same master copy
    sd = {k: v.to(torch.float16).to(torch.float32) for k, v in sd.items()}
    prompts = synth.make_prompts(B, seed=4321, ragged=ragged)
    g = torch.Generator().manual_seed(99)
    img = (torch.randn(B, 32, cfg.qformer_hidden, generator=g) * 0.5).float()
    dicom = [f"d{i}" for i in range(B)]
    model = R.build_ref_llama(cfg, sd, dtype, {d: img[i].numpy() for i, d in enumerate(dicom)})
    R.attach_lora(model, cfg, sd, dtype)
    t0 = time.time()
    with torch.no_grad():
        mask = prompts.ne(0).long()
        mi = model.prepare_inputs_for_generation(prompts, past_key_values=None, attention_mask=mask, use_cache=True, dicom=dicom)
        out = model(**mi, return_dict=True, output_hidden_states=True)
        ref_logits = out.logits
        ref_hidden = out.hidden_states
        ref_ids, ref_scores = R.ref_greedy(model, prompts, dicom, new_tokens, return_scores=True)
    t_ref = time.time() - t0

    orc = O.LlamaOracle(cfg, sd, dtype)
    pos = orc.positions_from_mask(mask)
    o_logits, _, o_hidden = orc.forward(prompts, mask, pos, None, img, return_hidden=True)
    o_ids, o_scores = orc.generate(prompts, img, new_tokens, return_scores=True)

    real = mask.bool()
    dl = (o_logits.float() - ref_logits.float()).abs()[real].max().item()
    dh = max((a.float() - b.float()).abs()[real].max().item() for a, b in zip(o_hidden, ref_hidden))
    same = bool(torch.equal(o_ids, ref_ids))
    print(f"[{tag}] ref {t_ref:.1f}s  oracle-vs-ref: max|dlogit|={dl:.3e} max|dhidden|={dh:.3e} "
          f"bit-equal logits={bool(torch.equal(o_logits[real], ref_logits[real]))} ids equal={same} "
          f"(len {ref_ids.shape[1]} vs {o_ids.shape[1]})")
    topv, topi = torch.stack([s.float() for s in ref_scores]).topk(8, dim=-1)
    np.savez_compressed(
        os.path.join(GOLD, f"llm_{tag}.npz"),
        cfg=np.array([cfg.vocab_size, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers,
                      cfg.num_attention_heads, cfg.max_position_embeddings]),
        dtype=dtype_name, seed=seed, img_seed=99, prompt_seed=4321, ragged=ragged, new_tokens=new_tokens,
        prompts=prompts.numpy(), sequences=ref_ids.numpy(),
        prefill_logits_last=_np(ref_logits[:, -1, :].float()).astype(np.float32),
        prefill_hidden_last_layer=_np(ref_hidden[-1][:, -4:, :].float()).astype(np.float32),
        step_topk_vals=topv.numpy().astype(np.float32), step_topk_idx=topi.numpy().astype(np.int32))
    return dl, same


def golden_vision(tag: str, B: int, image_size: int = 448):
    vcfg = synth.VisionCfg(image_size=image_size)
    sd = synth.make_vision_weights(vcfg, seed=0)
    imgs = synth.make_images(B, size=image_size, seed=1234)
    t0 = time.time()
    im = R.build_ref_image_model(sd)
    q_emb, q_enc = R.build_ref_qformer(sd, vcfg)
    ref_q, ref_e = R.ref_forward_image(im, q_emb, q_enc, sd, vcfg, imgs)
    t_ref = time.time() - t0
    o_q, o_e = O.forward_image(imgs, sd, vcfg)
    print(f"[{tag}] ref {t_ref:.1f}s oracle-vs-ref: max|dq|={(o_q - ref_q).abs().max():.3e} (|q|max {ref_q.abs().max():.3f}) "
          f"max|dembeds|={(o_e - ref_e).abs().max():.3e}")
    np.savez_compressed(os.path.join(GOLD, f"vision_{tag}.npz"), image_size=image_size, B=B, seed=0, img_seed=1234,
                        q_out=ref_q.numpy().astype(np.float32),
                        image_embeds_sub=ref_e[:, ::7, ::11].numpy().astype(np.float32))


def golden_vision_temporal(tag: str, B: int = 1, image_size: int = 448):
    """Two-image branch: the reference's own MultiImageEncoder + VisionTransformerPooler on (current, previous)."""
    import transformers.modeling_utils  # noqa: F401  (before the timm import shim: transformers probes timm's module spec)
    vcfg = synth.VisionCfg(image_size=image_size)
    sd = synth.make_vision_weights(vcfg, seed=0)
    cur = synth.make_images(B, size=image_size, seed=1234)
    prev = synth.make_images(B, size=image_size, seed=4242)
    im = R.build_ref_image_model(sd)
    q_emb, q_enc = R.build_ref_qformer(sd, vcfg)
    ref_q, ref_e = R.ref_forward_image(im, q_emb, q_enc, sd, vcfg, cur, prev)
    o_q, o_e = O.forward_image(cur, sd, vcfg, prev)
    single_q, _ = O.forward_image(cur, sd, vcfg)
    print(f"[{tag}] oracle-vs-ref: max|dq|={(o_q - ref_q).abs().max():.3e} (|q|max {ref_q.abs().max():.3f}) "
          f"max|dembeds|={(o_e - ref_e).abs().max():.3e}; two-image vs single-image max|dq|={(o_q - single_q).abs().max():.3e}")
    np.savez_compressed(os.path.join(GOLD, f"vision_{tag}.npz"), image_size=image_size, B=B, seed=0, img_seed=1234, prev_seed=4242,
                        q_out=ref_q.numpy().astype(np.float32), image_embeds_sub=ref_e[:, ::7, ::11].numpy().astype(np.float32))


def main():
    assert R.available(), "reference tree not found (this script only runs in the build container)"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    tiny = synth.tiny_llama_cfg()
    golden_llm("tiny_f32", tiny, "float32", B=3, new_tokens=6, ragged=True)
    golden_llm("tiny_f16", tiny, "float16", B=3, new_tokens=6, ragged=True)
    golden_llm("tiny_bf16", tiny, "bfloat16", B=3, new_tokens=6, ragged=True)
    # one real-width layer pair (H=4096, I=11008, 32 heads): the shapes the production kernels run
    wide = synth.LlamaCfg(num_hidden_layers=2)
    golden_llm("wide2_f16", wide, "float16", B=2, new_tokens=3, ragged=True)
    golden_vision("r50_448", B=2)
    golden_vision_temporal("temporal_r50_448", B=1)


if __name__ == "__main__":
    main()
