"""Host-side mirror of ``Blip2Qformer.forward_image`` (blip2_qformer.py:467-484) over the native vision engine.

``Blip2Qformer.from_state_dict(cfg, state_dict)`` takes tensors under the reference's own ``state_dict()`` names
(``visual_encoder.*``, ``ln_vision.*``, ``query_tokens``, ``Qformer.bert.*``), folds what is input independent on the
host (eval-mode BatchNorm into conv weight/bias, the constant ``missing_previous_emb`` half of the projector input into
its bias, the LayerNorm of the learned query tokens) and hands packed fp16/bf16 weights to ``libradialog_b200.so``.
``forward_image(image[B,3,S,S]) -> (q_out[B,32,768] fp32, image_embeds[B,196,1408] fp32)`` like the reference;
callers index ``[0]`` and ``.cpu().detach()`` (demo.py:270-273, pretraining/train.py:142-145).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from .synth import VisionCfg, _resnet_plan

STEM_KP = 152


def _fold_bn(w: torch.Tensor, sd, bn: str, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + eps)
    bias = sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale
    return w.float() * scale[:, None, None, None], bias


def _nhwc_rows(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] (the im2col column order of the engine)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def pack_vision_weights(cfg: VisionCfg, sd: Dict[str, torch.Tensor], dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    """name -> CPU tensor in the engine's layout; ``.w`` in ``dtype``, ``.b`` / ``.g`` fp32."""
    out: Dict[str, torch.Tensor] = {}
    R = "visual_encoder.encoder.encoder."
    w, b = _fold_bn(sd[R + "conv1.weight"], sd, R + "bn1", cfg.bn_eps)
    w = _nhwc_rows(w)
    out["conv1.w"] = F.pad(w, (0, STEM_KP - w.shape[1])).to(dtype)
    out["conv1.b"] = b
    for prefix, inplanes, planes, stride, down in _resnet_plan(cfg):
        p = R + prefix
        for ci in (1, 2, 3):
            w, b = _fold_bn(sd[f"{p}.conv{ci}.weight"], sd, f"{p}.bn{ci}", cfg.bn_eps)
            out[f"{prefix}.conv{ci}.w"] = _nhwc_rows(w).to(dtype)
            out[f"{prefix}.conv{ci}.b"] = b
        if down:
            w, b = _fold_bn(sd[p + ".downsample.0.weight"], sd, p + ".downsample.1", cfg.bn_eps)
            out[f"{prefix}.downsample.w"] = _nhwc_rows(w).to(dtype)
            out[f"{prefix}.downsample.b"] = b
    E, P = "visual_encoder.encoder.", "visual_encoder.projector.model."
    nb = cfg.backbone_to_vit
    out["b2v.w"] = _nhwc_rows(sd[E + "backbone_to_vit.weight"].float()).to(dtype)
    # projector conv1 sees cat([patch, missing_previous_emb]) (encoder.py:128-130): the second half is a constant
    w1 = sd[P + "0.weight"].float()[:, :, 0, 0]
    emb = sd[E + "missing_previous_emb"].float().reshape(-1)
    scale = sd[P + "1.weight"].float() / torch.sqrt(sd[P + "1.running_var"].float() + cfg.bn_eps)
    out["proj1.w"] = (w1[:, :nb] * scale[:, None]).contiguous().to(dtype)
    out["proj1.b"] = (w1[:, nb:] @ emb - sd[P + "1.running_mean"].float()) * scale + sd[P + "1.bias"].float()
    out["proj2.w"] = sd[P + "3.weight"].float()[:, :, 0, 0].contiguous().to(dtype)
    out["proj2.b"] = sd[P + "3.bias"].float()
    out["ln_vision.g"] = sd["ln_vision.weight"].float()
    out["ln_vision.b"] = sd["ln_vision.bias"].float()
    B = "Qformer.bert."
    # BertEmbeddings with input_ids=None is LayerNorm(query_tokens) (Qformer.py:78-108): input independent
    h0 = F.layer_norm(sd["query_tokens"].float()[0], (cfg.q_hidden,), sd[B + "embeddings.LayerNorm.weight"].float(),
                      sd[B + "embeddings.LayerNorm.bias"].float(), cfg.q_ln_eps)
    out["q.h0.w"] = h0.to(dtype)
    kv_w, kv_b = [], []
    for i in range(cfg.q_layers):
        p = B + f"encoder.layer.{i}."
        q = f"q{i}."
        out[q + "self_qkv.w"] = torch.cat([sd[p + f"attention.self.{n}.weight"].float() for n in ("query", "key", "value")], 0).to(dtype)
        out[q + "self_qkv.b"] = torch.cat([sd[p + f"attention.self.{n}.bias"].float() for n in ("query", "key", "value")], 0)
        out[q + "self_out.w"] = sd[p + "attention.output.dense.weight"].to(dtype)
        out[q + "self_out.b"] = sd[p + "attention.output.dense.bias"].float()
        out[q + "self_ln.g"] = sd[p + "attention.output.LayerNorm.weight"].float()
        out[q + "self_ln.b"] = sd[p + "attention.output.LayerNorm.bias"].float()
        if i % cfg.cross_attention_freq == 0:
            out[q + "cross_q.w"] = sd[p + "crossattention.self.query.weight"].to(dtype)
            out[q + "cross_q.b"] = sd[p + "crossattention.self.query.bias"].float()
            kv_w += [sd[p + "crossattention.self.key.weight"].float(), sd[p + "crossattention.self.value.weight"].float()]
            kv_b += [sd[p + "crossattention.self.key.bias"].float(), sd[p + "crossattention.self.value.bias"].float()]
            out[q + "cross_out.w"] = sd[p + "crossattention.output.dense.weight"].to(dtype)
            out[q + "cross_out.b"] = sd[p + "crossattention.output.dense.bias"].float()
            out[q + "cross_ln.g"] = sd[p + "crossattention.output.LayerNorm.weight"].float()
            out[q + "cross_ln.b"] = sd[p + "crossattention.output.LayerNorm.bias"].float()
        out[q + "ffn1.w"] = sd[p + "intermediate_query.dense.weight"].to(dtype)
        out[q + "ffn1.b"] = sd[p + "intermediate_query.dense.bias"].float()
        out[q + "ffn2.w"] = sd[p + "output_query.dense.weight"].to(dtype)
        out[q + "ffn2.b"] = sd[p + "output_query.dense.bias"].float()
        out[q + "ffn_ln.g"] = sd[p + "output_query.LayerNorm.weight"].float()
        out[q + "ffn_ln.b"] = sd[p + "output_query.LayerNorm.bias"].float()
    out["q.cross_kv.w"] = torch.cat(kv_w, 0).to(dtype)      # all cross-attention K/V projections as one GEMM
    out["q.cross_kv.b"] = torch.cat(kv_b, 0)
    # ---- two-image (temporal) branch: VisionTransformerPooler (biovil_t/transformer.py:28-118), optional in a state dict ----
    VP = E + "vit_pooler."
    if (VP + "norm_post.weight") in sd:
        C = nb
        for i in range(cfg.pooler_blocks):
            p, q = VP + f"blocks.{i}.", f"vp{i}."
            out[q + "ln1.g"], out[q + "ln1.b"] = sd[p + "norm1.weight"].float(), sd[p + "norm1.bias"].float()
            out[q + "qkv.w"] = torch.cat([sd[p + f"attn.{n}.weight"].float() for n in ("proj_q", "proj_k", "proj_v")], 0).to(dtype)
            out[q + "proj.w"], out[q + "proj.b"] = sd[p + "attn.proj.weight"].to(dtype), sd[p + "attn.proj.bias"].float()
            out[q + "ln2.g"], out[q + "ln2.b"] = sd[p + "norm2.weight"].float(), sd[p + "norm2.bias"].float()
            out[q + "fc1.w"], out[q + "fc1.b"] = sd[p + "mlp.fc1.weight"].to(dtype), sd[p + "mlp.fc1.bias"].float()
            out[q + "fc2.w"], out[q + "fc2.b"] = sd[p + "mlp.fc2.weight"].to(dtype), sd[p + "mlp.fc2.bias"].float()
        out["vp.post.g"], out["vp.post.b"] = sd[VP + "norm_post.weight"].float(), sd[VP + "norm_post.bias"].float()
        # pos_embed is a non-persistent buffer (transformer.py:63-65): recomputed; tokens = [current (type 0) ; previous (type 1)]
        pos = sine_position_embedding(cfg.grid, cfg.grid, C // 2)[0]
        te = sd[VP + "type_embed"].float()
        out["vp.pos_type.w"] = torch.cat([pos + te[0], pos + te[1]], 0).to(dtype)
        # projector conv1 over cat([patch, pooled difference tokens]): full [J, 2C] weight, BatchNorm folded
        out["proj1f.w"] = (w1 * scale[:, None]).contiguous().to(dtype)
        out["proj1f.b"] = (sd[P + "1.bias"].float() - sd[P + "1.running_mean"].float() * scale)
    return {k: v.contiguous() for k, v in out.items()}


def sine_position_embedding(H: int, W: int, embedding_dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """``SinePositionEmbedding(embedding_dim, normalize=True)(mask=ones[1,H,W])`` (biovil_t/transformer.py:225-266) -> [1, H*W, 2*dim]:
    normalised cumulative row / column indices scaled to 2*pi, interleaved sin / cos over ``temperature ** (2*(i//2)/dim)``."""
    import math
    ones = torch.ones(1, H, W)
    y, x = ones.cumsum(1, dtype=torch.float32), ones.cumsum(2, dtype=torch.float32)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(embedding_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / embedding_dim)
    px, py = x[:, :, :, None] / dim_t, y[:, :, :, None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).view(1, H * W, embedding_dim * 2)


class _IncompatibleKeys:
    """What ``nn.Module.load_state_dict(strict=False)`` returns (``load_checkpoint`` hands it back like base_model.py:52-56)."""

    def __init__(self, missing_keys, unexpected_keys):
        self.missing_keys, self.unexpected_keys = list(missing_keys), list(unexpected_keys)

    def __repr__(self):
        return f"_IncompatibleKeys(missing_keys={self.missing_keys}, unexpected_keys={self.unexpected_keys})"


def read_lavis_checkpoint(url_or_filename: str) -> Dict[str, torch.Tensor]:
    """LAVIS ``checkpoint_{epoch,best,last}.pth`` as written by ``RunnerBase._save_checkpoint`` (runner_base.py:658-683):
    ``{"model": state_dict WITHOUT the parameters that do not require grad, "optimizer", "config", "scaler", "epoch"}``;
    a bare state dict is accepted too (base_model.py:45-48).  No URL download: there is no network on this path."""
    import os
    if not os.path.isfile(url_or_filename):
        raise RuntimeError("checkpoint url or path is invalid")             # same message as base_model.py:43
    checkpoint = torch.load(url_or_filename, map_location="cpu", weights_only=False)
    return checkpoint["model"] if "model" in checkpoint.keys() else checkpoint


class Blip2Qformer:
    def __init__(self, cfg: VisionCfg, sd: Dict[str, torch.Tensor], torch_dtype: torch.dtype = torch.float16,
                 device="cuda:0", max_batch: int = 32):
        if not torch.cuda.is_available():
            raise RuntimeError("radialog_b200.Blip2Qformer needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = _lib.load()
        self.cfg = cfg
        self.dtype = torch_dtype
        self.device = torch.device(device)
        self._sd = {k: v.detach().cpu() for k, v in sd.items()}     # master copy under the reference's key names
        self._packed = {k: v.to(self.device) for k, v in pack_vision_weights(cfg, self._sd, torch_dtype).items()}
        self._h = None
        self._max_batch = 0
        self.reserve(max_batch)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._sd)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True) -> _IncompatibleKeys:
        """``nn.Module.load_state_dict`` semantics over the reference's key names; re-folds and re-packs the engine weights."""
        # keys of the reference module that this path never reads (text branch of the Q-Former, ITM/LM heads, temp) are not "unexpected"
        known = set(self._sd)
        missing = [k for k in self._sd if k not in state_dict]
        unexpected = [k for k in state_dict if k not in known]
        for k, v in state_dict.items():
            if k in known:
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise RuntimeError(f"size mismatch for {k}: copying a param with shape {tuple(v.shape)} from checkpoint, "
                                       f"the shape in current model is {tuple(self._sd[k].shape)}.")
                self._sd[k] = v.detach().cpu()
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]}... unexpected {unexpected[:5]}...")
        self._packed = {k: v.to(self.device) for k, v in pack_vision_weights(self.cfg, self._sd, self.dtype).items()}
        mb = self._max_batch
        self._destroy()
        self._max_batch = 0
        self.reserve(mb)
        return _IncompatibleKeys(missing, unexpected)

    def load_checkpoint(self, url_or_filename: str) -> _IncompatibleKeys:
        """``BaseModel.load_checkpoint`` (base_model.py:29-56) / ``Blip2Base.load_from_pretrained`` (blip2.py:88-104):
        a LAVIS checkpoint holds only the TRAINED parameters (the frozen image encoder and ``ln_vision`` are dropped by
        runner_base.py:665-671), loaded non-strictly over the live module."""
        return self.load_state_dict(read_lavis_checkpoint(url_or_filename), strict=False)

    load_from_pretrained = load_checkpoint

    @classmethod
    def from_state_dict(cls, cfg: VisionCfg, sd, **kw) -> "Blip2Qformer":
        return cls(cfg, sd, **kw)

    def to(self, *a, **k): return self       # demo.py:269,271 moves the model CPU<->GPU per image; weights stay resident here
    def eval(self): return self
    def cuda(self): return self

    def reserve(self, max_batch: int):
        if self._h is not None and max_batch <= self._max_batch:
            return
        self._destroy()
        c = self.cfg
        vc = _lib.VisionConfig(image_size=c.image_size, layers=(C.c_int * 4)(*c.layers), width=c.width, backbone_to_vit=c.backbone_to_vit,
                               joint=c.joint_feature_size, num_query=c.num_query_token, q_hidden=c.q_hidden, q_heads=c.q_heads,
                               q_layers=c.q_layers, q_inter=c.q_intermediate, cross_freq=c.cross_attention_freq,
                               ln_vision_eps=c.ln_vision_eps, q_ln_eps=c.q_ln_eps, dtype=_lib.dtype_code(self.dtype), max_batch=max_batch,
                               pooler_blocks=c.pooler_blocks if "vp.post.g" in self._packed else 0, pooler_heads=c.pooler_heads,
                               pooler_hidden=int(c.backbone_to_vit * c.pooler_mlp_ratio), pooler_ln_eps=c.pooler_ln_eps)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rd_vision_create(C.byref(vc), C.byref(h)), "rd_vision_create")
        self._h, self._max_batch = h, max_batch
        for name, t in self._packed.items():
            key = name[:-2] if name in ("q.h0.w", "vp.pos_type.w") else name
            _lib.check(self._lib.rd_vision_set_weight(h, key.encode(), _lib.ptr(t)), f"set_weight {name}")

    def _destroy(self):
        if self._h is not None:
            torch.cuda.synchronize()
            self._lib.rd_vision_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @torch.no_grad()
    def forward_image(self, image: torch.Tensor, previous_image: torch.Tensor = None):
        """``Blip2Qformer.forward_image`` (blip2_qformer.py:467-484).  ``previous_image`` (same shape) selects the two-image
        branch of the BioViL-T encoder (biovil_t/encoder.py:117-123: VisionTransformerPooler over current + previous tokens)
        - an extension: the reference's forward_image never passes a previous image."""
        c = self.cfg
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] != c.image_size or image.shape[3] != c.image_size:
            raise ValueError(f"image must be [B,3,{c.image_size},{c.image_size}], got {tuple(image.shape)}")
        if previous_image is not None:
            if previous_image.shape != image.shape:
                raise AssertionError("current_image and previous_image shapes do not match")     # encoder.py:118
            if "vp.post.g" not in self._packed:
                raise KeyError("the state dict holds no visual_encoder.encoder.vit_pooler.* weights: two-image mode is unavailable")
        img = image.to(self.device, torch.float32).contiguous()
        prev = None if previous_image is None else previous_image.to(self.device, torch.float32).contiguous()
        B = img.shape[0]
        q_out = torch.empty(B, c.num_query_token, c.q_hidden, device=self.device, dtype=torch.float32)
        embeds = torch.empty(B, c.num_patches, c.joint_feature_size, device=self.device, dtype=torch.float32)
        done = 0
        with torch.cuda.device(self.device):
            while done < B:      # chunk to the engine's reserved batch
                n = min(self._max_batch, B - done)
                if prev is None:
                    _lib.check(self._lib.rd_vision_forward(self._h, img[done:].data_ptr(), n, q_out[done:].data_ptr(), embeds[done:].data_ptr(),
                                                           _lib.current_stream()), "rd_vision_forward")
                else:
                    _lib.check(self._lib.rd_vision_forward_temporal(self._h, img[done:].data_ptr(), prev[done:].data_ptr(), n, q_out[done:].data_ptr(),
                                                                    embeds[done:].data_ptr(), _lib.current_stream()), "rd_vision_forward_temporal")
                done += n
        return q_out, embeds

    def launch_count(self) -> int:
        return int(self._lib.rd_vision_launch_count(self._h))
