"""ORACLE tooling — container-only.  Imports the UNMODIFIED reference modules from /root/reference so that
``make_golden.py`` can pin ``radialog_oracle.py`` against the reference itself.  Never imported by the product,
by ``-m gpu`` tests, by ``smoke()`` or by ``bench.py`` (/root/reference does not exist on the GPU box).

The shims below only satisfy import-time names of third-party packages that are absent or have moved
(SURVEY.md Appendix A); no reference arithmetic is replaced.
"""
from __future__ import annotations

import importlib.util
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("RADIALOG_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "biovil_t"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod          # transformers 5 looks the class module up in sys.modules
    spec.loader.exec_module(mod)
    return mod


_llama_mod = None


def llama_module():
    """modeling_llama_imgemb.py loaded by path (package import needs omegaconf)."""
    global _llama_mod
    if _llama_mod is None:
        _llama_mod = _load("ref_modeling_llama_imgemb",
                           os.path.join(REF, "model/lavis/models/blip2_models/modeling_llama_imgemb.py"))
    return _llama_mod


def build_ref_llama(cfg, sd, dtype, blip_embeddings: dict):
    """Construct the reference LlamaForCausalLM with our seeded weights.  The constructor reads
    pretraining/embs/..._test.pkl relative to CWD (modeling_llama_imgemb.py:461) -> run in a scratch dir."""
    from transformers import LlamaConfig
    mod = llama_module()
    scratch = tempfile.mkdtemp(prefix="rd_ref_")
    os.makedirs(os.path.join(scratch, "pretraining/embs"))
    with open(os.path.join(scratch, "pretraining/embs/stage1_pt_instruct_blip_origlr_img448_embeddings_test.pkl"), "wb") as f:
        pickle.dump({k: np.asarray(v, dtype=np.float32) for k, v in blip_embeddings.items()}, f)
    cwd = os.getcwd()
    os.chdir(scratch)
    try:
        hf = LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                         num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                         max_position_embeddings=cfg.max_position_embeddings, rms_norm_eps=cfg.rms_norm_eps,
                         pad_token_id=cfg.pad_token_id, bos_token_id=cfg.bos_token_id, eos_token_id=cfg.eos_token_id)
        model = mod.LlamaForCausalLM(hf)
    finally:
        os.chdir(cwd)
    model.model.img_proj_layer = nn.Linear(cfg.qformer_hidden, cfg.hidden_size)     # test.py:295
    base = {k: v for k, v in sd.items() if not k.startswith("base_model.")}
    missing, unexpected = model.load_state_dict(base, strict=False)
    missing = [m for m in missing if "rotary_emb" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    model = model.to(dtype).eval()
    return model


class _LoraLinear(nn.Module):
    """peft @ e536616 ``lora.Linear.forward`` (unmerged, eval): ``F.linear(x, W) + lora_B(lora_A(x)) * scaling``.
    peft is absent; this wrapper is the only restated arithmetic on the reference side."""

    def __init__(self, base: nn.Linear, A: torch.Tensor, B: torch.Tensor, scaling: float):
        super().__init__()
        self.base = base
        self.lora_A = nn.Linear(A.shape[1], A.shape[0], bias=False)
        self.lora_B = nn.Linear(B.shape[1], B.shape[0], bias=False)
        self.lora_A.weight.data = A.clone()
        self.lora_B.weight.data = B.clone()
        self.scaling = scaling

    def forward(self, x):
        result = self.base(x)
        result += self.lora_B(self.lora_A(x)) * self.scaling
        return result


def attach_lora(model, cfg, sd, dtype):
    for i, layer in enumerate(model.model.layers):
        for n in ("q_proj", "v_proj"):
            p = f"base_model.model.model.layers.{i}.self_attn.{n}."
            if p + "lora_A.weight" in sd:
                setattr(layer.self_attn, n, _LoraLinear(getattr(layer.self_attn, n), sd[p + "lora_A.weight"].to(dtype),
                                                        sd[p + "lora_B.weight"].to(dtype), cfg.lora_scaling))
    return model


@torch.no_grad()
def ref_greedy(model, input_ids, dicom, max_new_tokens, pad_id=0, eos_id=2, return_scores=False):
    """transformers 4.28.1 greedy_search restated around the reference's OWN prepare_inputs_for_generation +
    forward (``.generate`` no longer exists on PreTrainedModel)."""
    ids = input_ids.clone()
    mask = ids.ne(pad_id).long()
    unfinished = torch.ones(ids.shape[0], dtype=torch.long)
    past, scores = None, []
    for _ in range(max_new_tokens):
        mi = model.prepare_inputs_for_generation(ids, past_key_values=past, attention_mask=mask, use_cache=True, dicom=dicom)
        out = model(**mi, return_dict=True)
        logits = out.logits[:, -1, :]
        if return_scores:
            scores.append(logits.clone())
        tok = torch.argmax(logits, dim=-1)
        tok = tok * unfinished + pad_id * (1 - unfinished)
        ids = torch.cat([ids, tok[:, None]], dim=-1)
        mask = torch.cat([mask, mask.new_ones((mask.shape[0], 1))], dim=-1)
        past = out.past_key_values
        unfinished = unfinished.mul((tok != eos_id).long())
        if unfinished.max() == 0:
            break
    return (ids, scores) if return_scores else ids


# ------------------------------------------------------------------------------------------------
# BioViL-T + Q-Former
# ------------------------------------------------------------------------------------------------

def _install_vision_shims():
    if "timm.models.layers" not in sys.modules:
        timm = types.ModuleType("timm"); models = types.ModuleType("timm.models"); layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()

            def forward(self, x):
                return x

        class Mlp(nn.Module):
            def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
                super().__init__()
                out_features = out_features or in_features
                hidden_features = hidden_features or in_features
                self.fc1 = nn.Linear(in_features, hidden_features); self.act = act_layer()
                self.fc2 = nn.Linear(hidden_features, out_features); self.drop = nn.Dropout(drop)

            def forward(self, x):
                return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))

        layers.DropPath, layers.Mlp, layers.trunc_normal_ = DropPath, Mlp, torch.nn.init.trunc_normal_
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if "health_multimodal.common.device" not in sys.modules:
        hm = types.ModuleType("health_multimodal"); common = types.ModuleType("health_multimodal.common")
        device = types.ModuleType("health_multimodal.common.device")
        device.get_module_device = lambda m: next(m.parameters()).device
        hm.common, common.device = common, device
        sys.modules.update({"health_multimodal": hm, "health_multimodal.common": common,
                            "health_multimodal.common.device": device})
    import torchvision.models.resnet as tvr
    if not hasattr(tvr, "model_urls"):
        tvr.model_urls = {"resnet50": "", "resnet18": ""}


def build_ref_image_model(sd, full: bool = True):
    """The object blip2.py:82-84 builds: ImageModel(resnet50_multi_image, joint_feature_size=1408)."""
    _install_vision_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import biovil_t.resnet as bres
    bres.load_state_dict_from_url = lambda *a, **k: None       # no network (resnet.py:57-59)
    _orig = bres.ResNetHIML.load_state_dict
    bres.ResNetHIML.load_state_dict = lambda self, s, *a, **k: None if s is None else _orig(self, s, *a, **k)
    from biovil_t.model import ImageModel
    m = ImageModel(img_encoder_type="resnet50_multi_image", joint_feature_size=1408, pretrained_model_path=None)
    bres.ResNetHIML.load_state_dict = _orig
    own = {k[len("visual_encoder."):]: v for k, v in sd.items() if k.startswith("visual_encoder.")}
    missing, unexpected = m.load_state_dict(own, strict=False)
    missing = [k for k in missing if not (k.startswith("encoder.encoder.fc") or k.endswith("num_batches_tracked") or
                                           (k.startswith("encoder.vit_pooler") and not any(x.startswith("encoder.vit_pooler") for x in own)))]
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    return m.eval()


def build_ref_qformer(sd, vcfg):
    """BertEmbeddings + BertEncoder from the reference Qformer.py."""
    import transformers.modeling_utils as mu
    from transformers import pytorch_utils as pu
    for n in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, n):
            setattr(mu, n, getattr(pu, n))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), torch.tensor([]))
    qf = _load("ref_Qformer", os.path.join(REF, "model/lavis/models/blip2_models/Qformer.py"))
    from transformers import BertConfig
    c = BertConfig(hidden_size=vcfg.q_hidden, num_hidden_layers=vcfg.q_layers, num_attention_heads=vcfg.q_heads,
                   intermediate_size=vcfg.q_intermediate)
    c.encoder_width = vcfg.joint_feature_size
    c.add_cross_attention = True
    c.cross_attention_freq = vcfg.cross_attention_freq
    c.query_length = vcfg.num_query_token
    emb, enc = qf.BertEmbeddings(c), qf.BertEncoder(c)
    own_e = {k[len("Qformer.bert.embeddings."):]: v for k, v in sd.items() if k.startswith("Qformer.bert.embeddings.")}
    own_x = {k[len("Qformer.bert.encoder."):]: v for k, v in sd.items() if k.startswith("Qformer.bert.encoder.")}
    emb.load_state_dict(own_e, strict=False)
    missing, unexpected = enc.load_state_dict(own_x, strict=False)
    missing = [k for k in missing if (".intermediate.dense" not in k and ".output.dense" not in k.replace("attention.output", "x")
                                      and ".output.LayerNorm" not in k.replace("attention.output", "x"))]
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    return emb.eval(), enc.eval()


@torch.no_grad()
def ref_forward_image(image_model, q_emb, q_enc, sd, vcfg, image, previous_image=None):
    """blip2_qformer.py:467-484 driven through the imported reference modules (glue lines :469-484 and the
    fp32 LayerNorm subclass blip2.py:199-205 are the only restated lines).  With ``previous_image`` the reference's own
    MultiImageEncoder.forward runs its two-image branch (biovil_t/encoder.py:117-123 -> VisionTransformerPooler)."""
    if previous_image is None:
        proj = image_model(image).projected_patch_embeddings
    else:
        patch_x, pooled_x = image_model.encoder(image, previous_image=previous_image, return_patch_embeddings=True)
        proj = image_model.forward_post_encoder(patch_x, pooled_x).projected_patch_embeddings
    x = proj.reshape(image.shape[0], -1, vcfg.joint_feature_size)
    x = torch.nn.functional.layer_norm(x.float(), (vcfg.joint_feature_size,), sd["ln_vision.weight"], sd["ln_vision.bias"],
                                       vcfg.ln_vision_eps)
    B = image.shape[0]
    q = sd["query_tokens"].expand(B, -1, -1)
    h = q_emb(query_embeds=q)
    out = q_enc(h, attention_mask=torch.zeros(B, 1, 1, vcfg.num_query_token), head_mask=[None] * vcfg.q_layers,
                encoder_hidden_states=x, encoder_attention_mask=torch.zeros(B, 1, 1, x.shape[1]),
                query_length=vcfg.num_query_token, return_dict=True)
    return out.last_hidden_state, x
