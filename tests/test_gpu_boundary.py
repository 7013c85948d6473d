"""The loader surface the reference scripts actually call, replayed in their order against ``radialog_b200``.

* test.py:288-302  - ``LlamaForCausalLM.from_pretrained(dir, torch_dtype=float16, device_map='auto')`` -> set
  ``base_model.img_proj_layer`` -> ``resize_token_embeddings(len(tokenizer))`` -> ``.cuda()`` ->
  ``PeftModelForCausalLM.from_pretrained(model, dir, torch_dtype=float16, use_ram_optimized_load=False).half()`` ->
  ``.eval()`` -> ``generate(input_ids=..., dicom=..., return_dict_in_generate=True, output_scores=True, max_new_tokens=...)``.
* demo.py:225-236,269-297 - the same loader without the resize, image tokens through ``current_chat_img.pt`` (``use_img=True``).
* finetune.py:139-150 - what ``adapter_model.bin`` holds (peft keys + ``base_model.model.model.img_proj_layer.*``).
* runner_base.py:658-683 / base_model.py:29-56 - LAVIS ``checkpoint_*.pth`` for the Q-Former stage.

The on-disk inputs are synthetic (no network, no real checkpoints): an HF directory (``config.json`` + two
``pytorch_model-0000x-of-00002.bin`` shards + index), a peft directory (``adapter_config.json`` + ``adapter_model.bin``), the
embedding pickles relative to the CWD.  Expected ids come from the oracle on the same tensors.
"""
import json
import os
import pickle

import numpy as np
import pytest
import torch
import torch.nn as nn

from radialog_b200 import synth
from radialog_b200.llm import EMB_PKL_TEST, LlamaForCausalLM, PeftModelForCausalLM
from oracle import radialog_oracle as O
from parity_util import assert_ids_match

pytestmark = pytest.mark.gpu

V0 = 32000          # vocabulary of the base checkpoint; "<IMG>" becomes id 32000 after add_special_tokens + resize


def write_hf_dir(path, cfg, sd):
    """config.json + two weight shards + index, as `save_pretrained` of transformers 4.28 lays a LLaMA out."""
    os.makedirs(path, exist_ok=True)
    hc = {"architectures": ["LlamaForCausalLM"], "vocab_size": V0, "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
          "num_hidden_layers": cfg.num_hidden_layers, "num_attention_heads": cfg.num_attention_heads, "hidden_act": "silu",
          "max_position_embeddings": cfg.max_position_embeddings, "rms_norm_eps": cfg.rms_norm_eps, "pad_token_id": 0, "bos_token_id": 1,
          "eos_token_id": 2, "torch_dtype": "float16", "tie_word_embeddings": False}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hc, f)
    base = {k: v.half() for k, v in sd.items() if k.startswith("model.") and "img_proj" not in k or k == "lm_head.weight"}
    base["model.embed_tokens.weight"] = base["model.embed_tokens.weight"][:V0].clone()
    base["lm_head.weight"] = base["lm_head.weight"][:V0].clone()
    # HF checkpoints of that era also carry the (recomputable) rotary inv_freq buffers: must be ignored by the loader
    for i in range(cfg.num_hidden_layers):
        base[f"model.layers.{i}.self_attn.rotary_emb.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, cfg.head_dim, 2).float() / cfg.head_dim))
    keys = sorted(base)
    half = len(keys) // 2
    names = ["pytorch_model-00001-of-00002.bin", "pytorch_model-00002-of-00002.bin"]
    wm = {}
    for name, ks in zip(names, (keys[:half], keys[half:])):
        torch.save({k: base[k] for k in ks}, os.path.join(path, name))
        wm.update({k: name for k in ks})
    with open(os.path.join(path, "pytorch_model.bin.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": wm}, f)


def write_peft_dir(path, cfg, sd):
    """finetune.py:139-150: get_peft_model_state_dict keys (no adapter name) + the two img_proj_layer tensors."""
    os.makedirs(path, exist_ok=True)
    out = {k: v.half() for k, v in sd.items() if "lora_" in k}
    out["base_model.model.model.img_proj_layer.weight"] = sd["model.img_proj_layer.weight"].half()
    out["base_model.model.model.img_proj_layer.bias"] = sd["model.img_proj_layer.bias"].half()
    torch.save(out, os.path.join(path, "adapter_model.bin"))
    with open(os.path.join(path, "adapter_config.json"), "w") as f:
        json.dump({"peft_type": "LORA", "task_type": "CAUSAL_LM", "r": cfg.lora_r, "lora_alpha": cfg.lora_alpha, "lora_dropout": 0.1,
                   "target_modules": ["q_proj", "v_proj"], "bias": "none", "inference_mode": True, "base_model_name_or_path": "lmsys/vicuna-7b-v1.3"}, f)


@pytest.fixture()
def disk(tmp_path, monkeypatch):
    cfg = synth.tiny_llama_cfg(num_hidden_layers=2)
    sd = {k: v.half().float() for k, v in synth.make_llama_weights(cfg, seed=5, dtype=torch.float32).items()}
    write_hf_dir(str(tmp_path / "vicuna"), cfg, sd)
    write_peft_dir(str(tmp_path / "ckpt"), cfg, sd)
    g = torch.Generator().manual_seed(77)
    embs = {f"dicom{i}": (torch.randn(32, 768, generator=g) * 0.5).numpy().astype(np.float32) for i in range(3)}
    monkeypatch.chdir(tmp_path)
    os.makedirs(os.path.dirname(EMB_PKL_TEST), exist_ok=True)
    with open(EMB_PKL_TEST, "wb") as f:                    # modeling_llama_imgemb.py:461 (the train_all file is optional, :454-459)
        pickle.dump(embs, f)
    return cfg, sd, embs, tmp_path


def oracle_for(cfg, sd, new_row_embed, new_row_head):
    """The reference after resize_token_embeddings(32001): rows < 32000 from the checkpoint, row 32000 freshly initialised."""
    osd = dict(sd)
    osd["model.embed_tokens.weight"] = torch.cat([sd["model.embed_tokens.weight"][:V0], new_row_embed.float().cpu()], 0)
    osd["lm_head.weight"] = torch.cat([sd["lm_head.weight"][:V0], new_row_head.float().cpu()], 0)
    return O.LlamaOracle(cfg, osd, torch.float16)


def test_replay_of_test_py_loader_and_generate(cuda_dev, disk):
    cfg, sd, embs, tmp = disk
    # ---- test.py:288-302, verbatim call shapes ------------------------------------------------------------------------
    lang_model = LlamaForCausalLM.from_pretrained(str(tmp / "vicuna"), torch_dtype=torch.float16, device_map='auto')
    assert lang_model.config.vocab_size == V0 and len(lang_model.model.blip_embeddings) == 3
    lang_model.base_model.img_proj_layer = nn.Linear(768, lang_model.base_model.config.hidden_size).to(lang_model.base_model.device)
    lang_model.resize_token_embeddings(V0 + 1)
    assert lang_model.config.vocab_size == V0 + 1
    lang_model = lang_model.cuda()
    lang_model = PeftModelForCausalLM.from_pretrained(lang_model, str(tmp / "ckpt"), torch_dtype=torch.float16, use_ram_optimized_load=False).half()
    lang_model.eval()
    # the adapter's img_proj_layer replaced the freshly initialised one (finetune.py:141-143)
    assert torch.equal(lang_model.base_model.img_proj_layer.weight.detach().cpu().half(), sd["model.img_proj_layer.weight"].half())
    # ---- test.py:336-348 ------------------------------------------------------------------------------------------------
    prompts = synth.make_prompts(3, seed=4321, ragged=True)
    dicom_id = ["dicom2", "dicom0", "dicom1"]
    out = lang_model.generate(input_ids=prompts.to(cuda_dev), dicom=dicom_id, return_dict_in_generate=True, output_scores=True,
                              max_new_tokens=12)
    orc = oracle_for(cfg, sd, lang_model.model.w["embed"][V0:], lang_model.model.w["lm_head"][V0:])
    img = torch.tensor(np.array([embs[d] for d in dicom_id]))
    o_ids, o_scores = orc.generate(prompts, img, 12, return_scores=True)
    assert_ids_match(out.sequences.cpu(), o_ids, o_scores, prompts.shape[1], torch.float16, "test.py replay", min_exact_rows=1.0)
    scale = o_scores[0].float().abs().max().item()
    assert (out.scores[0].float().cpu() - o_scores[0].float()).abs().max().item() <= 1e-2 * scale
    with pytest.raises(KeyError):
        lang_model.generate(input_ids=prompts.to(cuda_dev), dicom=["nope", "dicom0", "dicom1"], max_new_tokens=2)


def test_replay_of_demo_py_loader_and_use_img(cuda_dev, disk):
    from radialog_b200 import pipeline
    cfg, sd, embs, tmp = disk
    # ---- demo.py:225-236 (no resize: "<IMG>" rows are never looked up because the splice overwrites them) ---------------
    lang_model = LlamaForCausalLM.from_pretrained(str(tmp / "vicuna"), torch_dtype=torch.float16, device_map='auto')
    lang_model.base_model.img_proj_layer = nn.Linear(768, lang_model.base_model.config.hidden_size).to(lang_model.base_model.device)
    lang_model = PeftModelForCausalLM.from_pretrained(lang_model, str(tmp / "ckpt"), torch_dtype=torch.float16, use_ram_optimized_load=False).half()
    lang_model.eval()
    # ---- demo.py:269-272: forward_image(...)[0] is [1,32,768]; saved with torch.save to the CWD ---------------------------
    qformer_embs = torch.from_numpy(embs["dicom1"])[None]
    pipeline.save_chat_image(qformer_embs)
    prompts = synth.make_prompts(1, seed=99)
    out = lang_model.generate(input_ids=prompts.to(cuda_dev), dicom=None, use_img=True, return_dict_in_generate=True, output_scores=True,
                              max_new_tokens=10)
    osd = dict(sd)
    osd["model.embed_tokens.weight"] = sd["model.embed_tokens.weight"][:V0]      # the 32 "<IMG>" ids are spliced, never looked up
    osd["lm_head.weight"] = sd["lm_head.weight"][:V0]                            # no resize in demo.py: 32000 logits
    orc = O.LlamaOracle(cfg, osd, torch.float16)
    o_ids, o_scores = orc.generate(prompts, qformer_embs, 10, return_scores=True)
    assert out.scores[0].shape[-1] == V0
    assert_ids_match(out.sequences.cpu(), o_ids, o_scores, prompts.shape[1], torch.float16, "demo.py replay", min_exact_rows=1.0)


def test_missing_files_fail_like_the_reference(cuda_dev, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    with pytest.raises(OSError):
        LlamaForCausalLM.from_pretrained(str(tmp_path / "nothing"), torch_dtype=torch.float16)
    cfg = synth.tiny_llama_cfg(num_hidden_layers=1)
    sd = synth.make_llama_weights(cfg, seed=1, dtype=torch.float32)
    write_hf_dir(str(tmp_path / "vicuna"), cfg, sd)
    with pytest.raises(FileNotFoundError):        # modeling_llama_imgemb.py:461: the test pickle is required
        LlamaForCausalLM.from_pretrained(str(tmp_path / "vicuna"), torch_dtype=torch.float16)
    m = LlamaForCausalLM.from_pretrained(str(tmp_path / "vicuna"), torch_dtype=torch.float16, load_embedding_pickles=False)
    with pytest.raises(AttributeError):           # img_proj_layer was never assigned (test.py:295 skipped)
        m.generate(input_ids=synth.make_prompts(1).to(cuda_dev), img_embeds=torch.zeros(1, 32, 768), max_new_tokens=2)


def test_lavis_checkpoint_round_trip(cuda_dev, tmp_path):
    """runner_base.py:658-683 writes {"model": trained params only, "optimizer", "config", "scaler", "epoch"}; base_model.py:29-56
    loads it non-strictly over the live module.  Loading a checkpoint of model B into an engine built from model A must give
    B's Q-Former on A's frozen image encoder - checked against the oracle on the merged state dict."""
    from radialog_b200.vision import Blip2Qformer
    cfg = synth.tiny_vision_cfg()
    sd_a = synth.make_vision_weights(cfg, seed=0)
    sd_b = synth.make_vision_weights(cfg, seed=7)
    frozen = ("visual_encoder.", "ln_vision.")                     # blip2_qformer.py:63-71 freezes these: dropped from checkpoints
    ckpt = {"model": {k: v for k, v in sd_b.items() if not k.startswith(frozen)},
            "optimizer": {"state": {}, "param_groups": []}, "config": {"run": {"task": "image_text_pretrain"}}, "scaler": None, "epoch": 3}
    ckpt["model"]["temp"] = torch.tensor(0.07)                     # a key of the reference module this path never reads
    path = str(tmp_path / "checkpoint_3.pth")
    torch.save(ckpt, path)
    model = Blip2Qformer.from_state_dict(cfg, sd_a, torch_dtype=torch.float16, device=cuda_dev, max_batch=2)
    imgs = synth.make_images(2, size=cfg.image_size, seed=1234)
    q_a, _ = model.forward_image(imgs.to(cuda_dev))
    msg = model.load_checkpoint(path)
    assert msg.unexpected_keys == ["temp"] and all(k.startswith(frozen) for k in msg.missing_keys) and msg.missing_keys
    q_b, _ = model.forward_image(imgs.to(cuda_dev))
    merged = dict(sd_a)
    merged.update({k: v for k, v in sd_b.items() if not k.startswith(frozen)})
    o_q, _ = O.forward_image(imgs, merged, cfg)
    rel = ((q_b.cpu() - o_q).abs().max() / o_q.abs().max()).item()
    assert rel <= 1e-2, f"after load_checkpoint: rel err {rel:.3e}"
    assert (q_a - q_b).abs().max().item() > 1e-3                   # the checkpoint really changed the Q-Former
    with pytest.raises(RuntimeError):
        model.load_checkpoint(str(tmp_path / "missing.pth"))
