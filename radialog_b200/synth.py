"""Seeded synthetic weights / inputs for the RaDialog image->report path.

There is no network and no checkpoint in this environment, so every weight the path
touches is generated from a seed.  Names follow the reference's ``state_dict()`` keys so
that a real checkpoint drops in unchanged:

* LLM: ``LlamaForCausalLM.state_dict()`` of
  model/lavis/models/blip2_models/modeling_llama_imgemb.py:675-701 plus
  ``model.img_proj_layer.{weight,bias}`` (test.py:295) and peft adapter keys
  ``base_model.model.model.layers.{i}.self_attn.{q,v}_proj.lora_{A,B}.weight``
  (finetune.py:167-173,311-318).
* Vision + Q-Former: ``Blip2Qformer.state_dict()`` keys (blip2_qformer.py:45-110):
  ``visual_encoder.encoder.encoder.*`` (torchvision ResNet-50), ``visual_encoder.encoder.backbone_to_vit``,
  ``visual_encoder.encoder.missing_previous_emb``, ``visual_encoder.projector.model.{0,1,3}``,
  ``ln_vision``, ``query_tokens``, ``Qformer.bert.*``.

Initialisers follow SURVEY.md section 8d (N(0,0.02) for LLaMA / BERT, kaiming for convs,
non-trivial BatchNorm running statistics so that BN folding is actually exercised).

Generation is pure ``torch`` CPU with an explicit ``torch.Generator`` and a fixed creation
order, so the same seed gives the same bits on every machine with the same torch build.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch

IMG_TOKEN_ID = 32000  # "<IMG>" placeholder id (test.py:296-297; modeling_llama_imgemb.py:500)
NUM_IMG_TOKENS = 32


@dataclass
class LlamaCfg:
    """Subset of HF ``LlamaConfig`` the path reads (Vicuna-7B-v1.3)."""
    vocab_size: int = 32001
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    max_position_embeddings: int = 2048
    rms_norm_eps: float = 1e-6
    pad_token_id: int = 0
    bos_token_id: int = 1
    eos_token_id: int = 2
    lora_r: int = 8          # finetune.py:167
    lora_alpha: int = 16     # finetune.py:168  -> scaling alpha/r = 2.0
    qformer_hidden: int = 768

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def lora_scaling(self) -> float:
        return self.lora_alpha / self.lora_r


def tiny_llama_cfg(**kw) -> LlamaCfg:
    """Small config with the real head_dim (128) used by unit-level parity tests."""
    base = dict(vocab_size=32001, hidden_size=256, intermediate_size=704, num_hidden_layers=2,
                num_attention_heads=2, max_position_embeddings=2048)
    base.update(kw)
    return LlamaCfg(**base)


@dataclass
class VisionCfg:
    """BioViL-T ResNet-50 + projector + Q-Former hyper-parameters."""
    image_size: int = 448
    layers: Tuple[int, ...] = (3, 4, 6, 3)       # biovil_t/resnet.py:80
    width: int = 64                               # torchvision ResNet base width
    backbone_to_vit: int = 256                    # biovil_t/encoder.py:95
    joint_feature_size: int = 1408                # blip2.py:82-84
    num_query_token: int = 32                     # blip2_pretrain_stage1_emb.yaml
    q_hidden: int = 768
    q_heads: int = 12
    q_layers: int = 12
    q_intermediate: int = 3072
    cross_attention_freq: int = 2                 # blip2.py:53
    pooler_blocks: int = 3                        # VisionTransformerPooler (biovil_t/transformer.py:41-44; encoder.py:104)
    pooler_heads: int = 8
    pooler_mlp_ratio: float = 1.0
    pooler_ln_eps: float = 1e-6                   # partial(nn.LayerNorm, eps=1e-6), transformer.py:45
    ln_vision_eps: float = 1e-5                   # nn.LayerNorm default (blip2.py:86)
    q_ln_eps: float = 1e-12                       # BertConfig.layer_norm_eps
    bn_eps: float = 1e-5

    @property
    def trunk_out(self) -> int:
        return self.width * 8 * 4

    @property
    def grid(self) -> int:
        return self.image_size // 32

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid


def tiny_vision_cfg(**kw) -> VisionCfg:
    """Reduced trunk/Q-Former for fast CPU tests; same op graph as the full model."""
    base = dict(image_size=64, layers=(1, 1, 1, 1), width=8, backbone_to_vit=32, joint_feature_size=64,
                num_query_token=32, q_hidden=64, q_heads=2, q_layers=2, q_intermediate=128, pooler_blocks=2, pooler_heads=1)
    base.update(kw)
    return VisionCfg(**base)


# --------------------------------------------------------------------------------------
# LLM weights
# --------------------------------------------------------------------------------------

def make_llama_weights(cfg: LlamaCfg, seed: int = 0, dtype: torch.dtype = torch.float16,
                       lora: bool = True, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """N(0, 0.02) Linear/Embedding, RMSNorm weight ~ 1 +- small (so the multiply is exercised),
    img_proj_layer default nn.Linear init, LoRA A kaiming-uniform and B ~ N(0,0.02)
    (not peft's zero init, so the side path is exercised).

    On ``device='cuda'`` the generator is a CUDA generator (used by bench.py for the 7B model,
    where CPU generation would take minutes); parity tests always use the CPU path.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    H, I, V, L = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers

    def normal(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std).to(dtype)

    def uniform(*shape, bound):
        return ((torch.rand(*shape, generator=g, device=device, dtype=torch.float32) * 2 - 1) * bound).to(dtype)

    sd: Dict[str, torch.Tensor] = {}
    emb = normal(V, H)
    emb[cfg.pad_token_id].zero_()  # nn.Embedding(padding_idx) zeroes that row (modeling_llama_imgemb.py:445)
    sd["model.embed_tokens.weight"] = emb
    for i in range(L):
        p = f"model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[p + f"self_attn.{n}.weight"] = normal(H, H)
        sd[p + "mlp.gate_proj.weight"] = normal(I, H)
        sd[p + "mlp.up_proj.weight"] = normal(I, H)
        sd[p + "mlp.down_proj.weight"] = normal(H, I)
        sd[p + "input_layernorm.weight"] = (1.0 + normal(H, std=0.05).float()).to(dtype)
        sd[p + "post_attention_layernorm.weight"] = (1.0 + normal(H, std=0.05).float()).to(dtype)
    sd["model.norm.weight"] = (1.0 + normal(H, std=0.05).float()).to(dtype)
    sd["lm_head.weight"] = normal(V, H)
    bound = 1.0 / math.sqrt(cfg.qformer_hidden)
    sd["model.img_proj_layer.weight"] = uniform(H, cfg.qformer_hidden, bound=bound)
    sd["model.img_proj_layer.bias"] = uniform(H, bound=bound)
    if lora:
        r = cfg.lora_r
        for i in range(L):
            for n in ("q_proj", "v_proj"):
                p = f"base_model.model.model.layers.{i}.self_attn.{n}."
                sd[p + "lora_A.weight"] = uniform(r, H, bound=math.sqrt(6.0 / ((1 + 5.0) * H)))  # kaiming_uniform(a=sqrt(5))
                sd[p + "lora_B.weight"] = normal(H, r)
    return sd


# --------------------------------------------------------------------------------------
# Vision + Q-Former weights (fp32 master copy, like the reference)
# --------------------------------------------------------------------------------------

def _resnet_plan(cfg: VisionCfg) -> List[Tuple[str, int, int, int, int]]:
    """(prefix, inplanes, planes, stride, has_downsample) per Bottleneck, torchvision order."""
    plan = []
    inplanes = cfg.width
    for li, nblocks in enumerate(cfg.layers):
        planes = cfg.width * (2 ** li)
        for b in range(nblocks):
            stride = 2 if (b == 0 and li > 0) else 1
            down = 1 if (b == 0 and (stride != 1 or inplanes != planes * 4)) else 0
            plan.append((f"layer{li + 1}.{b}", inplanes, planes, stride, down))
            inplanes = planes * 4
    return plan


def make_vision_weights(cfg: VisionCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed + 1000)
    sd: Dict[str, torch.Tensor] = {}

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def conv(name, cout, cin, k, bias=False):
        # kaiming_normal_(mode=fan_out, relu) as torchvision's ResNet does
        std = math.sqrt(2.0 / (cout * k * k))
        sd[name + ".weight"] = randn(cout, cin, k, k, std=std)
        if bias:
            sd[name + ".bias"] = randn(cout, std=0.02)

    def bn(name, c):
        # non-trivial affine + running statistics so BN folding is exercised (values stay O(1))
        sd[name + ".weight"] = 1.0 + randn(c, std=0.1)
        sd[name + ".bias"] = randn(c, std=0.1)
        sd[name + ".running_mean"] = randn(c, std=0.1)
        sd[name + ".running_var"] = 1.0 + 0.2 * torch.rand(c, generator=g, dtype=torch.float32)

    R = "visual_encoder.encoder.encoder."
    conv(R + "conv1", cfg.width, 3, 7)
    bn(R + "bn1", cfg.width)
    for prefix, inplanes, planes, stride, down in _resnet_plan(cfg):
        p = R + prefix
        conv(p + ".conv1", planes, inplanes, 1); bn(p + ".bn1", planes)
        conv(p + ".conv2", planes, planes, 3);   bn(p + ".bn2", planes)
        conv(p + ".conv3", planes * 4, planes, 1); bn(p + ".bn3", planes * 4)
        if down:
            conv(p + ".downsample.0", planes * 4, inplanes, 1); bn(p + ".downsample.1", planes * 4)
    E = "visual_encoder.encoder."
    conv(E + "backbone_to_vit", cfg.backbone_to_vit, cfg.trunk_out, 1)
    sd[E + "missing_previous_emb"] = torch.nn.init.trunc_normal_(
        torch.zeros(1, cfg.backbone_to_vit, 1, 1), std=0.02, generator=g)
    P = "visual_encoder.projector.model."
    J = cfg.joint_feature_size
    conv(P + "0", J, 2 * cfg.backbone_to_vit, 1)
    bn(P + "1", J)
    conv(P + "3", J, J, 1, bias=True)
    sd["ln_vision.weight"] = 1.0 + randn(J, std=0.05)
    sd["ln_vision.bias"] = randn(J, std=0.05)
    sd["query_tokens"] = randn(1, cfg.num_query_token, cfg.q_hidden, std=0.02)

    Hq, Iq = cfg.q_hidden, cfg.q_intermediate
    B = "Qformer.bert."

    def lin(name, nout, nin):
        sd[name + ".weight"] = randn(nout, nin, std=0.02)
        sd[name + ".bias"] = randn(nout, std=0.02)

    def ln(name, n):
        sd[name + ".weight"] = 1.0 + randn(n, std=0.05)
        sd[name + ".bias"] = randn(n, std=0.05)

    ln(B + "embeddings.LayerNorm", Hq)
    for i in range(cfg.q_layers):
        p = B + f"encoder.layer.{i}."
        for n in ("query", "key", "value"):
            lin(p + f"attention.self.{n}", Hq, Hq)
        lin(p + "attention.output.dense", Hq, Hq)
        ln(p + "attention.output.LayerNorm", Hq)
        if i % cfg.cross_attention_freq == 0:
            lin(p + "crossattention.self.query", Hq, Hq)
            lin(p + "crossattention.self.key", Hq, J)
            lin(p + "crossattention.self.value", Hq, J)
            lin(p + "crossattention.output.dense", Hq, Hq)
            ln(p + "crossattention.output.LayerNorm", Hq)
        lin(p + "intermediate_query.dense", Iq, Hq)
        lin(p + "output_query.dense", Hq, Iq)
        ln(p + "output_query.LayerNorm", Hq)
    # VisionTransformerPooler of the two-image (temporal) branch, biovil_t/transformer.py:28-75 / encoder.py:104: 3 blocks of
    # dim = backbone_to_vit, 8 heads, mlp_ratio 1, q/k/v without bias.  Drawn LAST so that every tensor above keeps the bits
    # it had before this branch existed (the committed fixtures depend on them).
    C = cfg.backbone_to_vit
    VP = E + "vit_pooler."
    for i in range(cfg.pooler_blocks):
        p = VP + f"blocks.{i}."
        ln(p + "norm1", C)
        for n in ("proj_q", "proj_k", "proj_v"):
            sd[p + f"attn.{n}.weight"] = randn(C, C, std=0.02)
        lin(p + "attn.proj", C, C)
        ln(p + "norm2", C)
        lin(p + "mlp.fc1", int(C * cfg.pooler_mlp_ratio), C)
        lin(p + "mlp.fc2", C, int(C * cfg.pooler_mlp_ratio))
    ln(VP + "norm_post", C)
    sd[VP + "type_embed"] = torch.nn.init.trunc_normal_(torch.zeros(2, 1, C), std=0.02, generator=g)
    return sd


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------

def make_images(n: int, size: int = 448, seed: int = 1234) -> torch.Tensor:
    """``ToTensor`` + ``ExpandChannels`` look-alike: 3 identical channels in [0,1) (ReportDataset.py:80-106)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 1, size, size, generator=g).repeat(1, 3, 1, 1).contiguous()


def make_prompts(n: int, seed: int = 4321, ragged: bool = False, vocab: int = 32000,
                 n_left: int = 15, n_right: int = 16) -> torch.Tensor:
    """Rows ``[bos] + n_left ids + [<IMG>]*32 + n_right ids`` (T=64 by default).  ``ragged`` draws the
    text lengths per row and left-pads with 0 (tokenizer.padding_side='left', pad=unk=0: test.py:291,304)."""
    g = torch.Generator().manual_seed(seed)
    T = 1 + n_left + NUM_IMG_TOKENS + n_right
    rows = []
    for _ in range(n):
        nl, nr = n_left, n_right
        if ragged:
            nl = int(torch.randint(max(1, n_left - 6), n_left + 1, (1,), generator=g))
            nr = int(torch.randint(max(1, n_right - 6), n_right + 1, (1,), generator=g))
        left = torch.randint(3, vocab, (nl,), generator=g)
        right = torch.randint(3, vocab, (nr,), generator=g)
        row = torch.cat([torch.tensor([1]), left, torch.full((NUM_IMG_TOKENS,), IMG_TOKEN_ID), right])
        rows.append(torch.cat([torch.zeros(T - row.numel(), dtype=torch.long), row]))
    return torch.stack(rows).long()
