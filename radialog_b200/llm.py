"""Host-side mirror of the reference's LLM surface for the image->report path.

Drop-in for ``model.lavis.models.blip2_models.modeling_llama_imgemb.LlamaForCausalLM`` as used by ``test.py:288-348``
and ``demo.py:221-297``: same attribute names (``.model`` / ``.base_model`` with a settable ``img_proj_layer``,
``.config.hidden_size``, ``.device``), ``resize_token_embeddings``, ``.cuda()/.half()/.eval()``, a peft-like
``PeftModelForCausalLM.from_pretrained`` and ``generate(input_ids, dicom=..., use_img=..., max_new_tokens=...,
return_dict_in_generate=True, output_scores=True)`` returning ``.sequences`` / ``.scores`` with the semantics of
transformers 4.28.1 greedy search.  All compute happens in ``libradialog_b200.so`` (sm_100a CUDA); PyTorch only owns
device memory and streams.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .synth import IMG_TOKEN_ID, NUM_IMG_TOKENS, LlamaCfg

EMB_PKL_TRAIN = "pretraining/embs/stage1_pt_instruct_blip_origlr_img448_embeddings_train_all.pkl"
EMB_PKL_TEST = "pretraining/embs/stage1_pt_instruct_blip_origlr_img448_embeddings_test.pkl"
CHAT_IMG_FILE = "current_chat_img.pt"


class _Config:
    """The attributes of HF ``LlamaConfig`` that callers read (``config.hidden_size``: test.py:295, demo.py:229)."""

    def __init__(self, cfg: LlamaCfg):
        self.vocab_size = cfg.vocab_size
        self.hidden_size = cfg.hidden_size
        self.intermediate_size = cfg.intermediate_size
        self.num_hidden_layers = cfg.num_hidden_layers
        self.num_attention_heads = cfg.num_attention_heads
        self.max_position_embeddings = cfg.max_position_embeddings
        self.rms_norm_eps = cfg.rms_norm_eps
        self.pad_token_id = cfg.pad_token_id
        self.bos_token_id = cfg.bos_token_id
        self.eos_token_id = cfg.eos_token_id


@dataclass
class GreedySearchDecoderOnlyOutput:
    """``.sequences`` [B, T+n] (prompt included, pad after EOS) and ``.scores`` (tuple of n [B,V] tensors)."""
    sequences: torch.Tensor
    scores: Optional[Tuple[torch.Tensor, ...]] = None


@dataclass
class BeamSearchDecoderOnlyOutput:
    """``generate(num_beams=k, return_dict_in_generate=True)``: best hypothesis per row, its length-normalised score, and (with
    ``output_scores``) the per-step log-probabilities of all B*k beams."""
    sequences: torch.Tensor
    sequences_scores: Optional[torch.Tensor] = None
    scores: Optional[Tuple[torch.Tensor, ...]] = None


class _BeamHypotheses:
    """n-best list of finished hypotheses of one batch item (transformers 4.28.1 ``BeamHypotheses``: score = sum of log-probs /
    len(hypothesis, prompt included) ** length_penalty; ``is_done`` compares the worst kept score with the best still attainable)."""

    def __init__(self, num_beams, length_penalty, early_stopping, max_length):
        self.num_beams, self.length_penalty, self.early_stopping, self.max_length = num_beams, length_penalty, early_stopping, max_length
        self.beams: List[Tuple[float, torch.Tensor]] = []
        self.worst_score = 1e9

    def add(self, hyp: torch.Tensor, sum_logprobs: float):
        score = sum_logprobs / (hyp.shape[-1] ** self.length_penalty)
        if len(self.beams) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self.beams) > self.num_beams:
                ranked = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[ranked[0][1]]
                self.worst_score = ranked[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        if len(self.beams) < self.num_beams:
            return False
        if self.early_stopping is True:
            return True
        ref_len = self.max_length if (self.early_stopping not in (True, False) and self.length_penalty > 0.0) else cur_len
        return self.worst_score >= best_sum_logprobs / ref_len ** self.length_penalty


def rope_tables(head_dim: int, max_pos: int, base: float = 10000.0):
    """LlamaRotaryEmbedding.__init__ (modeling_llama_imgemb.py:97-109): fp32 tables built on the host."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_pos, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


class LlamaModel:
    """``lang_model.model`` / ``lang_model.base_model``: holds the packed weights and the image-embedding side channels
    (modeling_llama_imgemb.py:433-462)."""

    def __init__(self, cfg: LlamaCfg, dtype: torch.dtype, device: torch.device, load_embedding_pickles: bool = False):
        self.cfg = cfg
        self.config = _Config(cfg)
        self.dtype = dtype
        self.device = device
        self.img_proj_layer: Optional[nn.Linear] = None       # set by callers exactly like test.py:295 / demo.py:229
        self.blip_embeddings: Dict[str, np.ndarray] = {}
        if load_embedding_pickles:                              # reference behaviour (:454-462): train optional, test required
            import pickle
            try:
                with open(EMB_PKL_TRAIN, "rb") as f:
                    self.blip_embeddings = pickle.load(f)
            except Exception:
                self.blip_embeddings = {}
                print("WARNING: no train blip embeddings found! For inference that is ok, if you want to train a new model, "
                      "please generate them as described in the readme.")
            with open(EMB_PKL_TEST, "rb") as f:                 # FileNotFoundError propagates like the reference
                self.blip_embeddings.update(pickle.load(f))
        self.w: Dict[str, torch.Tensor] = {}
        self.layers_w: List[Dict[str, torch.Tensor]] = []
        self.has_lora = False


class LlamaForCausalLM:
    # ------------------------------------------------------------------------------------------
    # construction
    # ------------------------------------------------------------------------------------------
    def __init__(self, cfg: LlamaCfg, torch_dtype: torch.dtype = torch.float16, device: Union[str, torch.device] = "cuda:0",
                 load_embedding_pickles: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("radialog_b200.LlamaForCausalLM needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = _lib.load()
        self.cfg = cfg
        self.config = _Config(cfg)
        self.dtype = torch_dtype
        self.device = torch.device(device)
        self.model = LlamaModel(cfg, torch_dtype, self.device, load_embedding_pickles)
        self._h = None
        self._cap = (0, 0)
        self._graphs: Dict[Tuple[int, bool], torch.cuda.CUDAGraph] = {}
        self._graph_kernels: Dict[Tuple[int, bool], int] = {}
        self._flag_host: Optional[torch.Tensor] = None
        self._timing_events = None
        self._capture_launches = 0
        self._replayed_kernels = 0
        self._n_prompt_cached = 0
        self._cached_ids: Optional[torch.Tensor] = None     # device copy of the ids whose KV are in the cache
        self._img_w_packed = None
        self.algo = _lib.ALGO_AUTO
        self.qkv_partials = True  # QKV GEMM hands its fp32 split-K partials to the attention kernel (default)
        self.od_partials = True   # o_proj / down_proj partials finished by the norm launch that follows them (default)
        self.use_cuda_graph = True
        self.last_stats: Dict[str, float] = {}

    @property
    def base_model(self):
        return self.model

    # no-ops kept for call-site compatibility (test.py:299-302, demo.py:236)
    def cuda(self, *a, **k): return self
    def half(self): return self
    def eval(self): return self
    def to(self, *a, **k): return self

    @classmethod
    def from_state_dict(cls, cfg: LlamaCfg, sd: Dict[str, torch.Tensor], torch_dtype=torch.float16, device="cuda:0",
                        **kw) -> "LlamaForCausalLM":
        m = cls(cfg, torch_dtype, device, **kw)
        m.load_state_dict(sd)
        return m

    @classmethod
    def from_pretrained(cls, path: str, torch_dtype=torch.float16, device_map=None, device="cuda:0", **kw) -> "LlamaForCausalLM":
        """Loads an HF-format LLaMA directory (config.json + pytorch_model*.bin shards), test.py:289 / demo.py:225."""
        cfg_path = os.path.join(path, "config.json")
        if not os.path.exists(cfg_path):
            raise OSError(f"{path} does not contain config.json (no network access: weights must be on disk)")
        with open(cfg_path) as f:
            hc = json.load(f)
        cfg = LlamaCfg(vocab_size=hc["vocab_size"], hidden_size=hc["hidden_size"], intermediate_size=hc["intermediate_size"],
                       num_hidden_layers=hc["num_hidden_layers"], num_attention_heads=hc["num_attention_heads"],
                       max_position_embeddings=hc.get("max_position_embeddings", 2048), rms_norm_eps=hc.get("rms_norm_eps", 1e-6),
                       pad_token_id=hc.get("pad_token_id", 0) or 0, bos_token_id=hc.get("bos_token_id", 1), eos_token_id=hc.get("eos_token_id", 2))
        sd: Dict[str, torch.Tensor] = {}
        for fn in sorted(os.listdir(path)):
            if fn.startswith("pytorch_model") and fn.endswith(".bin"):
                sd.update(torch.load(os.path.join(path, fn), map_location="cpu"))
        if not sd:
            raise OSError(f"no pytorch_model*.bin shards under {path}")
        return cls.from_state_dict(cfg, sd, torch_dtype, device, load_embedding_pickles=kw.get("load_embedding_pickles", True))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Packs reference-named tensors into the engine layout: fused q|k|v, fused gate|up, LoRA kept unmerged."""
        cfg, dt, dev = self.cfg, self.dtype, self.device
        H, r = cfg.hidden_size, cfg.lora_r

        def g(name):
            return sd[name].to(device=dev, dtype=dt).contiguous()

        m = self.model
        m.w["embed"] = g("model.embed_tokens.weight")
        m.w["final_norm"] = g("model.norm.weight")
        m.w["lm_head"] = g("lm_head.weight")
        cos, sin = rope_tables(cfg.head_dim, cfg.max_position_embeddings)
        m.w["cos"], m.w["sin"] = cos.to(dev, dt).contiguous(), sin.to(dev, dt).contiguous()    # :123-124 cast to x.dtype
        lora_prefix = "base_model.model.model.layers.{}.self_attn.{}."
        m.has_lora = (lora_prefix.format(0, "q_proj") + "lora_A.weight") in sd
        m.layers_w = []
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}."
            lw = {
                "qkv": torch.cat([g(p + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")], 0).contiguous(),
                "o": g(p + "self_attn.o_proj.weight"),
                "gate_up": torch.cat([g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")], 0).contiguous(),
                "down": g(p + "mlp.down_proj.weight"),
                "ln1": g(p + "input_layernorm.weight"),
                "ln2": g(p + "post_attention_layernorm.weight"),
            }
            if m.has_lora:
                aq, av = g(lora_prefix.format(i, "q_proj") + "lora_A.weight"), g(lora_prefix.format(i, "v_proj") + "lora_A.weight")
                bq, bv = g(lora_prefix.format(i, "q_proj") + "lora_B.weight"), g(lora_prefix.format(i, "v_proj") + "lora_B.weight")
                self._pack_lora(lw, aq, av, bq, bv)
            m.layers_w.append(lw)
        if "model.img_proj_layer.weight" in sd:
            lin = nn.Linear(cfg.qformer_hidden, H)
            lin.weight.data = sd["model.img_proj_layer.weight"].float().clone()
            lin.bias.data = sd["model.img_proj_layer.bias"].float().clone()
            m.img_proj_layer = lin.to(dev)
        self._destroy_engine()
        return self

    def _pack_lora(self, lw, aq, av, bq, bv):
        """Unmerged LoRA in the engine layout: the two lora_A blocks ride as 2r extra rows of the fused QKV weight (the GEMM
        then also produces t = lora_A . x), lora_B [2H, r] is applied where q and v are consumed."""
        H3 = 3 * self.cfg.hidden_size
        lw["qkv"] = torch.cat([lw["qkv"][:H3], aq, av], 0).contiguous()
        lw["lora_b"] = torch.cat([bq, bv], 0).contiguous()
        lw.pop("lora_a", None)

    def load_adapter(self, adapter_sd: Dict[str, torch.Tensor]):
        """peft adapter_model.bin: lora_A/lora_B of q_proj,v_proj and img_proj_layer (finetune.py:139-145)."""
        merged = {}
        for k, v in adapter_sd.items():
            merged[k] = v
            if k.endswith("img_proj_layer.weight"):
                merged["model.img_proj_layer.weight"] = v
            if k.endswith("img_proj_layer.bias"):
                merged["model.img_proj_layer.bias"] = v
        cfg, dt, dev, H, r = self.cfg, self.dtype, self.device, self.cfg.hidden_size, self.cfg.lora_r
        m = self.model
        for i, lw in enumerate(m.layers_w):
            p = f"base_model.model.model.layers.{i}.self_attn."
            self._pack_lora(lw, merged[p + "q_proj.lora_A.weight"].to(dev, dt), merged[p + "v_proj.lora_A.weight"].to(dev, dt),
                            merged[p + "q_proj.lora_B.weight"].to(dev, dt), merged[p + "v_proj.lora_B.weight"].to(dev, dt))
        m.has_lora = True
        if "model.img_proj_layer.weight" in merged:
            lin = nn.Linear(cfg.qformer_hidden, H)
            lin.weight.data = merged["model.img_proj_layer.weight"].float().clone()
            lin.bias.data = merged["model.img_proj_layer.bias"].float().clone()
            m.img_proj_layer = lin.to(dev)
        self._destroy_engine()
        return self

    def resize_token_embeddings(self, n: int):
        """HF semantics (test.py:297): grow/shrink embed_tokens and lm_head; new rows N(0, 0.02)."""
        m = self.model
        old = m.w["embed"].shape[0]
        if n == old:
            return
        for key in ("embed", "lm_head"):
            w = m.w[key]
            new = torch.zeros(n, w.shape[1], device=w.device, dtype=w.dtype)
            k = min(n, old)
            new[:k] = w[:k]
            if n > old:
                new[old:] = (torch.randn(n - old, w.shape[1], device=w.device) * 0.02).to(w.dtype)
            m.w[key] = new.contiguous()
        self.cfg.vocab_size = n
        self.config.vocab_size = n
        m.config.vocab_size = n
        self._destroy_engine()

    # ------------------------------------------------------------------------------------------
    # engine plumbing
    # ------------------------------------------------------------------------------------------
    def _destroy_engine(self):
        if self._h is not None:
            torch.cuda.synchronize()
            self._lib.rd_llm_destroy(self._h)
            self._h = None
        self._graphs = {}
        self._graph_kernels = {}
        self._capture_launches = 0
        self._replayed_kernels = 0
        self._cap = (0, 0)
        self._cached_ids = None

    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:
            pass

    def reserve(self, max_batch: int, max_ctx: int):
        """(Re)creates the native engine with room for ``max_batch`` rows of ``max_ctx`` tokens (KV cache + scratch)."""
        if self._h is not None and max_batch <= self._cap[0] and max_ctx <= self._cap[1]:
            return
        max_batch = max(max_batch, self._cap[0])
        max_ctx = min(max(max_ctx, self._cap[1]), self.cfg.max_position_embeddings)
        self._destroy_engine()
        cfg, m = self.cfg, self.model
        c = _lib.LlmConfig(vocab=cfg.vocab_size, hidden=cfg.hidden_size, inter=cfg.intermediate_size, layers=cfg.num_hidden_layers,
                           heads=cfg.num_attention_heads, max_pos=cfg.max_position_embeddings, rms_eps=cfg.rms_norm_eps,
                           dtype=_lib.dtype_code(self.dtype), lora_r=cfg.lora_r if m.has_lora else 0, lora_scale=cfg.lora_scaling,
                           qformer_hidden=cfg.qformer_hidden, max_batch=max_batch, max_ctx=max_ctx, pad_id=cfg.pad_token_id,
                           eos_id=cfg.eos_token_id, img_id=IMG_TOKEN_ID)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rd_llm_create(C.byref(c), C.byref(h)), "rd_llm_create")
        self._h = h
        self._cap = (max_batch, max_ctx)
        sw = self._lib.rd_llm_set_weight
        for slot, key in ((_lib.W_EMBED, "embed"), (_lib.W_FINAL_NORM, "final_norm"), (_lib.W_LM_HEAD, "lm_head"),
                          (_lib.W_ROPE_COS, "cos"), (_lib.W_ROPE_SIN, "sin")):
            _lib.check(sw(h, -1, slot, _lib.ptr(m.w[key])), f"set_weight {key}")
        slots = {"qkv": _lib.W_QKV, "o": _lib.W_O, "gate_up": _lib.W_GATE_UP, "down": _lib.W_DOWN, "ln1": _lib.W_LN1,
                 "ln2": _lib.W_LN2, "lora_a": _lib.W_LORA_A, "lora_b": _lib.W_LORA_B}
        for i, lw in enumerate(m.layers_w):
            for key, t in lw.items():
                _lib.check(sw(h, i, slots[key], _lib.ptr(t)), f"set_weight layer {i} {key}")
        _lib.check(self._lib.rd_llm_set_algo(h, self.algo), "set_algo")
        _lib.check(self._lib.rd_llm_set_qkv_partials(h, 1 if self.qkv_partials else 0), "set_qkv_partials")
        _lib.check(self._lib.rd_llm_set_od_partials(h, 1 if self.od_partials else 0), "set_od_partials")

    def _bind_img_proj(self):
        lin = self.model.img_proj_layer
        if lin is None:
            raise AttributeError("img_proj_layer is not set (callers assign nn.Linear(768, hidden_size): test.py:295)")
        w = lin.weight.detach().to(self.device, self.dtype).contiguous()
        # bias is rounded to the storage dtype first (the reference's Linear holds it in fp16 after .half()), kept as fp32 for the epilogue
        b = lin.bias.detach().to(self.device, self.dtype).float().contiguous()
        self._img_w_packed = (w, b)
        _lib.check(self._lib.rd_llm_set_weight(self._h, -1, _lib.W_IMG_PROJ_W, _lib.ptr(w)), "set img_proj w")
        _lib.check(self._lib.rd_llm_set_weight(self._h, -1, _lib.W_IMG_PROJ_B, _lib.ptr(b)), "set img_proj b")

    def set_algo(self, algo: int):
        self.algo = algo
        self._graphs = {}
        if self._h is not None:
            _lib.check(self._lib.rd_llm_set_algo(self._h, algo), "set_algo")

    def set_od_partials(self, on: bool):
        """Single-token steps: o_proj / down_proj leave fp32 split-K partials and the norm launch that follows sums them, adds the
        residual and normalises (default), or the GEMMs reduce over their cluster and plain norm kernels follow."""
        self.od_partials = bool(on)
        self._graphs = {}
        if self._h is not None:
            _lib.check(self._lib.rd_llm_set_od_partials(self._h, 1 if on else 0), "set_od_partials")

    def set_qkv_partials(self, on: bool):
        """Single-token steps: the QKV GEMM leaves fp32 split-K partials for the attention kernel to sum (default), or reduces
        them itself over a thread-block cluster."""
        self.qkv_partials = bool(on)
        self._graphs = {}
        if self._h is not None:
            _lib.check(self._lib.rd_llm_set_qkv_partials(self._h, 1 if on else 0), "set_qkv_partials")

    def _state(self):
        gen, fin, logits, hidden = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_int()
        _lib.check(self._lib.rd_llm_state(self._h, C.byref(gen), C.byref(fin), C.byref(logits), C.byref(hidden), C.byref(n)), "state")
        return gen.value, fin.value, logits.value, hidden.value, n.value

    def _wrap(self, ptr: int, shape, dtype):
        """Zero-copy torch view over engine-owned device memory (read-only use)."""
        typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.float16: "<f2", torch.uint8: "|u1", torch.bfloat16: "<i2"}[dtype]
        iface = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        holder = type("_Dev", (), {"__cuda_array_interface__": iface})()
        t = torch.as_tensor(holder, device=self.device)
        return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t

    # ------------------------------------------------------------------------------------------
    # forward for parity dumps: full-position logits like LlamaForCausalLM.forward (:705-793)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prefill_logits(self, input_ids: torch.Tensor, img_embeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        ids = input_ids.to(self.device).long().contiguous()
        B, T = ids.shape
        self.reserve(B, T + 2)
        img = self._prep_img(ids, img_embeds)
        out = torch.empty(B, T, self.cfg.vocab_size, device=self.device, dtype=self.dtype)
        _lib.check(self._lib.rd_llm_prefill(self._h, _lib.ptr(ids), _lib.ptr(img), B, T, _lib.ptr(out), 0, _lib.current_stream()), "prefill")
        torch.cuda.synchronize()
        self._cached_ids = None
        return out

    def _prep_img(self, ids: torch.Tensor, img_embeds: Optional[torch.Tensor]):
        if img_embeds is None:
            return None
        B = ids.shape[0]
        img = img_embeds
        if img.dim() == 2:
            img = img[None]
        if img.shape[0] == 1 and B > 1:
            img = img.expand(B, -1, -1)
        if tuple(img.shape) != (B, NUM_IMG_TOKENS, self.cfg.qformer_hidden):
            raise ValueError(f"image embeddings must be [{B},{NUM_IMG_TOKENS},{self.cfg.qformer_hidden}], got {tuple(img.shape)}")
        cnt = (ids == IMG_TOKEN_ID).sum(-1)
        if not bool(((cnt == 0) | (cnt == NUM_IMG_TOKENS)).all()):
            raise ValueError("each row must contain either no <IMG> token or one run of exactly 32 (split_at_img, "
                             "modeling_llama_imgemb.py:498-520)")
        self._bind_img_proj()
        return img.to(self.device, self.dtype).contiguous()       # .half() cast of the Q-Former output (:576/:579)

    # ------------------------------------------------------------------------------------------
    # generate
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor = None, dicom: Optional[Sequence[str]] = None, use_img: bool = False,
                 return_dict_in_generate: bool = False, output_scores: bool = False, max_new_tokens: int = 20, num_beams: int = 1,
                 img_embeds: Optional[torch.Tensor] = None, suppress_eos: bool = False, reuse_cache: bool = False,
                 check_every: int = 1, forced_tokens: Optional[torch.Tensor] = None, **unused):
        """Greedy decoding with the semantics of transformers 4.28.1 ``generate`` as the reference calls it
        (test.py:339-348, demo.py:290-297): attention mask inferred as ``ids != pad``, left padding, EOS rows emit pad.

        ``dicom`` looks the Q-Former tokens up in ``self.model.blip_embeddings`` (KeyError if unknown, like the reference);
        ``use_img`` reads ``current_chat_img.pt`` from the CWD; ``img_embeds`` is the direct device-tensor hand-off.
        ``reuse_cache`` keeps the KV cache of the previous call and only runs the new suffix (multi-turn chat).
        ``num_beams > 1`` is HF beam search (test.py:267,467,629) over the flat KV cache, see ``_beam_search``.
        ``forced_tokens`` [B, n] (parity harness): step s consumes ``forced_tokens[:, s]`` instead of the engine's own
        choice (teacher forcing); ``.sequences`` still holds the engine's choices, ``.scores`` its logits."""
        if input_ids is None:
            raise ValueError("You have to specify either decoder_input_ids or decoder_inputs_embeds")
        ids = input_ids.to(self.device).long().contiguous()
        if ids.dim() != 2:
            raise ValueError(f"input_ids must be [B,T], got {tuple(ids.shape)}")
        B, T = ids.shape
        if img_embeds is None:
            if use_img:
                img_embeds = torch.load(CHAT_IMG_FILE)                                       # modeling_llama_imgemb.py:576
            elif dicom is not None:
                img_embeds = torch.tensor(np.array([self.model.blip_embeddings[d] for d in dicom]))   # :579 (KeyError propagates)
        if T + max_new_tokens + 1 > self.cfg.max_position_embeddings:
            raise ValueError(f"prompt ({T}) + max_new_tokens ({max_new_tokens}) exceeds max_position_embeddings")
        if num_beams != 1:
            if reuse_cache or forced_tokens is not None:
                raise ValueError("beam search does not combine with reuse_cache / forced_tokens")
            return self._beam_search(ids, img_embeds, num_beams, max_new_tokens, return_dict_in_generate, output_scores,
                                     float(unused.get("length_penalty", 1.0)), bool(unused.get("early_stopping", False)))
        if forced_tokens is not None:
            forced_tokens = forced_tokens.to(self.device).long()
            if forced_tokens.shape[0] != B or forced_tokens.shape[1] < max_new_tokens - 1:
                raise ValueError("forced_tokens must be [B, >= max_new_tokens-1]")
            forced_cols = [forced_tokens[:, s].contiguous() for s in range(forced_tokens.shape[1])]
        need_ctx = T + max_new_tokens + 1
        if not (reuse_cache and self._h is not None and B <= self._cap[0] and need_ctx <= self._cap[1]):
            self.reserve(B, max(need_ctx, 128))
        with torch.cuda.device(self.device):
            return self._generate_greedy(ids, img_embeds, B, T, max_new_tokens, suppress_eos, reuse_cache, check_every,
                                         forced_cols if forced_tokens is not None else None, return_dict_in_generate, output_scores)

    def _generate_greedy(self, ids, img_embeds, B, T, max_new_tokens, suppress_eos, reuse_cache, check_every, forced_cols,
                         return_dict_in_generate, output_scores):
        st = _lib.current_stream()
        ev0, ev1, ev2 = self._events()
        ev0.record()

        # ---- prefill (or suffix-only extend when the cached conversation is a prefix of the new one) --------------------
        start = 0
        if reuse_cache and self._cached_ids is not None and self._cached_ids.shape[0] == B:
            start = self._common_prefix(ids)                  # device-side compare, one small read-back (multi-turn only)
        if start > 0:
            npos = (ids[:, :start] != self.cfg.pad_token_id).sum(-1).to(torch.int32).cpu().contiguous()
            _lib.check(self._lib.rd_llm_truncate(self._h, start, npos.data_ptr(), st), "truncate")
            suffix = ids[:, start:].contiguous()
            _lib.check(self._lib.rd_llm_extend(self._h, _lib.ptr(suffix), B, T - start, int(suppress_eos), st), "extend")
        else:
            img = self._prep_img(ids, img_embeds)
            _lib.check(self._lib.rd_llm_prefill(self._h, _lib.ptr(ids), _lib.ptr(img), B, T, None, int(suppress_eos), st), "prefill")
        ev1.record()
        gen_p, fin_p, logits_p, _, _ = self._state()
        Cmax, V = self._cap[1], self.cfg.vocab_size
        vpad = (V + 63) // 64 * 64
        gen_t = self._wrap(gen_p, (self._cap[0], Cmax), torch.int64)
        logits_t = self._wrap(logits_p, (self._cap[0], vpad), self.dtype)
        scores = []
        if output_scores:
            scores.append(logits_t[:B, :V].clone())

        # ---- decode loop: one native call (or one CUDA-graph replay) per token.  The stop rule of greedy_search ("all rows
        # ---- finished") is a device word that is copied to pinned memory after every step and read WITHOUT blocking, so the
        # ---- host keeps enqueueing ahead of the device and overshoots the real end by the few steps it is ahead.
        flag_p = C.c_void_p()
        _lib.check(self._lib.rd_llm_done_flag(self._h, C.byref(flag_p)), "done_flag")
        flag_t = self._wrap(flag_p.value, (1,), torch.int32)
        poll = not suppress_eos
        if poll:
            if self._flag_host is None or self._flag_host.numel() < max_new_tokens + 1:
                self._flag_host = torch.zeros(max(max_new_tokens + 1, 512), dtype=torch.int32).pin_memory()
            flags = self._flag_host
            flags[:max_new_tokens + 1].zero_()
            flags[0:1].copy_(flag_t, non_blocking=True)          # after the prefill's selection
            poll_events = [torch.cuda.Event()]
            poll_events[0].record()
            oldest = 0
        n_done = 1
        stop_at = 0
        while n_done < max_new_tokens and not stop_at:
            if forced_cols is not None:
                _lib.check(self._lib.rd_llm_force_tokens(self._h, _lib.ptr(forced_cols[n_done - 1]), st), "force_tokens")
            self._decode_one(B, st, n_done, suppress_eos)
            n_done += 1
            if output_scores:
                scores.append(logits_t[:B, :V].clone())
            if poll and (n_done % check_every == 0 or n_done == max_new_tokens):
                flags[n_done - 1:n_done].copy_(flag_t, non_blocking=True)
                e = torch.cuda.Event()
                e.record()
                poll_events.append(e)
                while oldest < len(poll_events) and poll_events[oldest].query():
                    oldest += 1
                done = flags[:n_done].max().item() if oldest > 0 else 0      # host memory: no device synchronisation
                if done:
                    stop_at = int(done)
        ev2.record()
        torch.cuda.synchronize(self.device)
        gen = gen_t[:B, :n_done].clone()
        n_keep = n_done
        if not suppress_eos:
            # HF stops right after the step at which the last unfinished row emitted EOS
            is_eos = gen == self.cfg.eos_token_id
            has = is_eos.any(-1)
            if bool(has.all()):
                first = torch.where(is_eos, torch.arange(n_done, device=gen.device)[None], n_done).min(-1).values
                n_keep = int(first.max().item()) + 1
        gen = gen[:, :n_keep]
        sequences = torch.cat([ids, gen], dim=-1)
        # ids whose K/V are in the cache now: prompt + all tokens that were fed back (device tensor; no host copy)
        fed = gen_t[:B, :n_done - 1] if forced_cols is None else torch.stack(forced_cols[:n_done - 1], 1) if n_done > 1 else gen_t[:B, :0]
        self._cached_ids = torch.cat([ids, fed], dim=-1)
        self._n_prompt_cached = T
        self.last_stats = {"prefill_ms": ev0.elapsed_time(ev1), "decode_ms": ev1.elapsed_time(ev2), "new_tokens": n_done,
                           "prefill_tokens": T - start, "reused_tokens": start}
        if return_dict_in_generate:
            return GreedySearchDecoderOnlyOutput(sequences=sequences, scores=tuple(scores[:n_keep]) if output_scores else None)
        return sequences

    def _beam_search(self, ids, img_embeds, num_beams, max_new_tokens, return_dict_in_generate, output_scores, length_penalty,
                     early_stopping):
        """``generate(num_beams=k)`` (test.py:467,629): transformers 4.28.1 ``beam_search`` + ``BeamSearchScorer`` semantics
        (num_return_sequences = 1, no logits processors).  The model forward, the cache reordering (``_reorder_cache``,
        modeling_llama_imgemb.py:838-843 -> ``rd_llm_reorder_cache``) and the token hand-over run in the native engine; the
        scorer's bookkeeping over the 2k candidates per row is host logic like in transformers.  Image rows are expanded with
        the prompts (the reference's ``dicom`` list is not expanded by transformers, so its own forward fails for k > 1)."""
        B, T = ids.shape
        nb, V = int(num_beams), self.cfg.vocab_size
        pad, eos = self.cfg.pad_token_id, self.cfg.eos_token_id
        max_length = T + max_new_tokens
        ids_x = ids.repeat_interleave(nb, dim=0).contiguous()                       # _expand_inputs_for_generation
        if img_embeds is not None:
            img = img_embeds if img_embeds.dim() == 3 else img_embeds[None]
            if img.shape[0] == 1 and B > 1:
                img = img.expand(B, -1, -1)
            img_embeds = img.repeat_interleave(nb, dim=0)
        BB = B * nb
        self.reserve(BB, max(max_length + 1, 128))
        self._graphs = {}
        st = _lib.current_stream()
        with torch.cuda.device(self.device):
            img = self._prep_img(ids_x, img_embeds)
            _lib.check(self._lib.rd_llm_prefill(self._h, _lib.ptr(ids_x), _lib.ptr(img), BB, T, None, 0, st), "prefill")
            _, _, logits_p, _, _ = self._state()
            vpad = (V + 63) // 64 * 64
            logits_t = self._wrap(logits_p, (self._cap[0], vpad), self.dtype)
            hyps = [_BeamHypotheses(nb, length_penalty, early_stopping, max_length) for _ in range(B)]
            done = [False] * B
            beam_scores = torch.zeros((B, nb), dtype=torch.float32, device=self.device)
            beam_scores[:, 1:] = -1e9
            beam_scores = beam_scores.view(-1)
            seqs = ids_x
            scores_out = []
            while True:
                step_scores = torch.log_softmax(logits_t[:BB, :V], dim=-1)            # in the logits dtype, like transformers
                if output_scores:
                    scores_out.append(step_scores.clone())
                cand = step_scores + beam_scores[:, None]                             # float32 by type promotion
                # transformers uses torch.topk, which leaves the order of EQUAL scores unspecified; a stable descending sort pins
                # it to "lowest flat index first" (one valid instance of that behaviour, reproducible across devices)
                srt_scores, srt_tokens = torch.sort(cand.view(B, nb * V), dim=1, descending=True, stable=True)
                top_scores, top_tokens = srt_scores[:, : 2 * nb], srt_tokens[:, : 2 * nb]
                top_idx = torch.div(top_tokens, V, rounding_mode="floor")
                top_tokens = top_tokens % V
                # ---- BeamSearchScorer.process on the host: B x 2k candidates ------------------------------------------
                ts, tt, ti = top_scores.cpu(), top_tokens.cpu(), top_idx.cpu()
                seqs_cpu = None
                cur_len = seqs.shape[-1]
                nxt_scores = torch.zeros((B, nb), dtype=torch.float32)
                nxt_tokens = torch.zeros((B, nb), dtype=torch.int64)
                nxt_idx = torch.zeros((B, nb), dtype=torch.int64)
                for b in range(B):
                    if done[b]:
                        nxt_tokens[b, :] = pad
                        continue
                    k = 0
                    for rank in range(2 * nb):
                        tok, row = int(tt[b, rank]), b * nb + int(ti[b, rank])
                        if tok == eos:
                            if rank >= nb:
                                continue
                            if seqs_cpu is None:
                                seqs_cpu = seqs.cpu()
                            hyps[b].add(seqs_cpu[row].clone(), float(ts[b, rank]))
                        else:
                            nxt_scores[b, k], nxt_tokens[b, k], nxt_idx[b, k] = ts[b, rank], tok, row
                            k += 1
                        if k == nb:
                            break
                    if k < nb:
                        raise ValueError(f"At most {nb} tokens in {tt[b].tolist()} can be equal to `eos_token_id: {eos}`. "
                                         "Make sure they are trained correctly.")
                    done[b] = done[b] or hyps[b].is_done(float(ts[b].max()), cur_len)
                beam_scores = nxt_scores.view(-1).to(self.device)
                beam_tokens = nxt_tokens.view(-1).to(self.device)
                beam_idx = nxt_idx.view(-1).to(self.device)
                seqs = torch.cat([seqs[beam_idx, :], beam_tokens[:, None]], dim=-1)
                if all(done) or seqs.shape[-1] >= max_length:
                    break
                _lib.check(self._lib.rd_llm_reorder_cache(self._h, _lib.ptr(beam_idx.to(torch.int32).contiguous()), st), "reorder_cache")
                _lib.check(self._lib.rd_llm_force_tokens(self._h, _lib.ptr(beam_tokens.contiguous()), st), "force_tokens")
                _lib.check(self._lib.rd_llm_decode_step(self._h, st), "decode_step")
            # ---- BeamSearchScorer.finalize ---------------------------------------------------------------------------
            seqs_cpu, bs_cpu = seqs.cpu(), beam_scores.cpu()
            for b in range(B):
                if not done[b]:
                    for k in range(nb):
                        hyps[b].add(seqs_cpu[b * nb + k], float(bs_cpu[b * nb + k]))
            best, best_scores = [], torch.zeros(B, dtype=torch.float32)
            for b in range(B):
                sc, hyp = sorted(hyps[b].beams, key=lambda x: x[0]).pop()
                best.append(hyp)
                best_scores[b] = sc
            lens = [int(x.shape[-1]) for x in best]
            sent_max_len = min(max(lens) + 1, max_length)
            decoded = torch.full((B, sent_max_len), pad, dtype=torch.int64)
            for b, hyp in enumerate(best):
                decoded[b, : lens[b]] = hyp
                if lens[b] < sent_max_len:
                    decoded[b, lens[b]] = eos
        self._graphs = {}
        self._cached_ids = None
        sequences = decoded.to(self.device)
        if return_dict_in_generate:
            return BeamSearchDecoderOnlyOutput(sequences=sequences, sequences_scores=best_scores.to(self.device),
                                               scores=tuple(scores_out) if output_scores else None)
        return sequences

    def _events(self):
        if self._timing_events is None:
            self._timing_events = tuple(torch.cuda.Event(enable_timing=True) for _ in range(3))
        return self._timing_events

    def _decode_one(self, B: int, st: int, n_done: int, suppress_eos: bool = False):
        if not self.use_cuda_graph or n_done < 2:
            # the first decode step always runs eagerly (lazy function-attribute setup must not happen under capture)
            _lib.check(self._lib.rd_llm_decode_step(self._h, st), "decode_step")
            return
        # a captured step bakes in its by-value kernel arguments (suppress_eos of the selection kernel): one graph per flag value
        key = (B, bool(suppress_eos))
        g = self._graphs.get(key)
        if g is None:
            before = int(self._lib.rd_llm_launch_count(self._h))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                _lib.check(self._lib.rd_llm_decode_step(self._h, _lib.current_stream()), "decode_step(capture)")
            _lib.check(self._lib.rd_llm_note_replayed_steps(self._h, -1), "note")   # capture recorded, did not run
            self._graphs[key] = g
            self._graph_kernels[key] = int(self._lib.rd_llm_launch_count(self._h)) - before
            self._capture_launches += self._graph_kernels[key]
        g.replay()
        self._replayed_kernels += self._graph_kernels[key]
        _lib.check(self._lib.rd_llm_note_replayed_steps(self._h, 1), "note")

    def launch_count(self) -> int:
        """Kernels of libradialog_b200 launched so far (eager launches + kernels inside replayed CUDA graphs)."""
        if self._h is None:
            return 0
        return int(self._lib.rd_llm_launch_count(self._h)) - self._capture_launches + self._replayed_kernels

    @torch.no_grad()
    def profile_decode_steps(self, B: int, T: int, steps: int = 8) -> Dict[str, Dict[str, float]]:
        """Per-kernel-class device time of `steps` eager decode steps (CUDA events around every launch) at context ~T."""
        from .synth import make_prompts
        ids = make_prompts(B, seed=1)[:, :T].to(self.device).contiguous()
        self.reserve(B, T + steps + 4)
        st = _lib.current_stream()
        _lib.check(self._lib.rd_llm_prefill(self._h, _lib.ptr(ids), None, B, T, None, 1, st), "prefill")
        _lib.check(self._lib.rd_llm_decode_step(self._h, st), "decode_step")       # warm
        torch.cuda.synchronize()
        _lib.check(self._lib.rd_llm_profile(self._h, 1), "profile on")
        for _ in range(steps):
            _lib.check(self._lib.rd_llm_decode_step(self._h, st), "decode_step")
        n = len(_lib.PROFILE_CLASSES)
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        _lib.check(self._lib.rd_llm_profile_read(self._h, ms, cnt, n), "profile read")
        _lib.check(self._lib.rd_llm_profile(self._h, 0), "profile off")
        self._cached_ids = None
        return {name: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, name in enumerate(_lib.PROFILE_CLASSES)}

    def _common_prefix(self, ids: torch.Tensor) -> int:
        """Longest prefix (same for all rows) of the new conversation whose KV entries are already cached and valid.
        Runs on the device tensors; only the two resulting scalars are read back."""
        cached = self._cached_ids
        L = min(cached.shape[1], ids.shape[1] - 1)       # always leave one token to run
        if L <= 0:
            return 0
        eq = cached[:, :L] == ids[:, :L]
        # a generated pad(0) has mask 1 in the cache but would be masked by a fresh prefill: stop before it
        gen_region = torch.arange(L, device=ids.device)[None] >= self._n_prompt_cached
        ok = eq & ~(gen_region & (ids[:, :L] == self.cfg.pad_token_id))
        bad = (~ok).float().cumsum(-1) > 0
        per_row = (~bad).sum(-1)
        # the <IMG> block must be entirely inside the reused prefix (extend does not splice)
        img_cols = (ids == IMG_TOKEN_ID).any(0)
        last_img = torch.where(img_cols, torch.arange(ids.shape[1], device=ids.device), -1).max()
        start, last_img = (int(v) for v in torch.stack([per_row.min(), last_img]).tolist())
        if start <= last_img:
            return 0
        return start


class PeftModelForCausalLM:
    """Shim with the call shape of ``peft.PeftModelForCausalLM.from_pretrained(model, path, torch_dtype=...,
    use_ram_optimized_load=False)`` (test.py:301, demo.py:232-234): loads adapter_model.bin into the wrapped model
    (LoRA stays unmerged, like peft in eval) and returns it."""

    @staticmethod
    def from_pretrained(model: LlamaForCausalLM, path_or_state_dict, torch_dtype=None, use_ram_optimized_load=False, **kw):
        if isinstance(path_or_state_dict, dict):
            sd = path_or_state_dict
        else:
            fn = os.path.join(path_or_state_dict, "adapter_model.bin")
            sd = torch.load(fn, map_location="cpu")
            cfg_fn = os.path.join(path_or_state_dict, "adapter_config.json")
            if os.path.exists(cfg_fn):
                with open(cfg_fn) as f:
                    ac = json.load(f)
                model.cfg.lora_r = ac.get("r", model.cfg.lora_r)
                model.cfg.lora_alpha = ac.get("lora_alpha", model.cfg.lora_alpha)
        return model.load_adapter(sd)
