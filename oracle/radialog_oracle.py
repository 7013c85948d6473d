"""ORACLE — test infrastructure only.  NOT part of the product path.

CPU (torch) restatement of the RaDialog image->report inference path, written from the
reference's algorithm, each function citing the reference file:line it follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / the CPU baseline.

Parity status: **pinned against the reference itself run in the build container**.  The reference
has no tests, golden vectors or fixtures for this path, so
``oracle/make_golden.py`` imports the reference's own modules from /root/reference
(modeling_llama_imgemb.py, Qformer.py leaf modules, biovil_t/*), loads the same seeded
weights (radialog_b200/synth.py), runs them, and commits inputs' seeds + the reference's outputs
under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this file against those fixtures.

Numerics conventions
--------------------
* ``dtype=torch.float32``: plain fp32 everywhere (what the reference computes on CPU).
* ``dtype=torch.float16`` / ``bfloat16``: every elementwise op is evaluated as torch evaluates
  it on half tensors (compute in fp32, round once per op), every matmul accumulates in fp32
  and rounds once — the behaviour of the reference's fp16 ``nn.Linear`` / ``torch.matmul`` on
  cuBLAS.  This reproduces the reference's rounding points listed in SURVEY.md Appendix B.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IMG_TOKEN_ID = 32000
NUM_IMG_TOKENS = 32


# ======================================================================================
# LLM (modeling_llama_imgemb.py)
# ======================================================================================

def _mm(x: torch.Tensor, w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """x @ w.T with fp32 accumulation, one rounding to ``dtype`` (fp16 nn.Linear without bias)."""
    return (x.float() @ w.float().t()).to(dtype)


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm.forward, modeling_llama_imgemb.py:85-93: fp32 variance, ``x * rsqrt`` in fp32,
    round to the weight dtype, THEN multiply by the weight in that dtype."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    h = x * torch.rsqrt(var + eps)
    if w.dtype in (torch.float16, torch.bfloat16):
        h = h.to(w.dtype)
    return w * h


def rope_tables(head_dim: int, max_pos: int, base: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """LlamaRotaryEmbedding.__init__, modeling_llama_imgemb.py:97-109 (fp32 tables [max_pos, head_dim])."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(max_pos, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    """modeling_llama_imgemb.py:128-132."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q, k, cos, sin, position_ids):
    """apply_rotary_pos_emb, modeling_llama_imgemb.py:135-142.  q,k: [B,nh,T,hd]; cos/sin [max_pos,hd] already
    cast to q.dtype (:123-124); three separately rounded ops per tensor."""
    c = cos[position_ids][:, None, :, :]
    s = sin[position_ids][:, None, :, :]
    return (q * c) + (_rotate_half(q) * s), (k * c) + (_rotate_half(k) * s)


def make_attention_mask(attn_mask_2d: torch.Tensor, q_len: int, dtype: torch.dtype) -> torch.Tensor:
    """_prepare_decoder_attention_mask / _make_causal_mask / _expand_mask, modeling_llama_imgemb.py:44-73,475-496.
    attn_mask_2d: [B, c] of {0,1}.  Returns additive [B,1,q_len,c] in ``dtype`` (may contain -inf where
    causal and padding masks add; re-clamped in ``attention``)."""
    B, c = attn_mask_2d.shape
    past = c - q_len
    minv = torch.finfo(dtype).min
    expanded = attn_mask_2d[:, None, None, :].expand(B, 1, q_len, c).to(dtype)
    inverted = 1.0 - expanded
    pad = inverted.masked_fill(inverted.to(torch.bool), minv)
    if q_len > 1:
        m = torch.full((q_len, q_len), minv, dtype=torch.float32)
        cond = torch.arange(q_len)
        m.masked_fill_(cond < (cond + 1).view(q_len, 1), 0)
        m = m.to(dtype)
        if past > 0:
            m = torch.cat([torch.zeros(q_len, past, dtype=dtype), m], dim=-1)
        return pad + m[None, None]
    return pad


def attention(q, k, v, mask, dtype) -> torch.Tensor:
    """LlamaAttention.forward core, modeling_llama_imgemb.py:216-234.  q [B,nh,q,hd], k/v [B,nh,c,hd].
    scores = (q k^T) rounded, THEN / sqrt(hd) rounded, + mask, max(., finfo.min), softmax in fp32
    rounded to ``dtype``, then @ v."""
    hd = q.shape[-1]
    w = (q.float() @ k.float().transpose(2, 3)).to(dtype) / math.sqrt(hd)
    w = w + mask
    w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min, dtype=w.dtype))
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(dtype)
    return (w.float() @ v.float()).to(dtype)


class LlamaOracle:
    """Functional restatement of LlamaForCausalLM (+ unmerged peft LoRA on q_proj / v_proj)."""

    def __init__(self, cfg, sd: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float16, use_lora: bool = True):
        self.cfg = cfg
        self.dtype = dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
        self.use_lora = use_lora and any("lora_A" in k for k in sd)
        cos, sin = rope_tables(cfg.head_dim, cfg.max_position_embeddings)
        self.cos, self.sin = cos.to(dtype), sin.to(dtype)   # :123-124 cast to x.dtype

    # -- pieces -------------------------------------------------------------------------
    def _proj(self, x, layer: int, name: str):
        """fp16 Linear; for q_proj/v_proj the unmerged peft LoRA forward ``Wx + (B(A x)) * scaling``
        (peft @ e536616 lora.Linear.forward; finetune.py:167-173): A, B are fp16 Linears, dropout is
        identity in eval, ``* scaling`` and ``+=`` are each rounded."""
        w = self.sd[f"model.layers.{layer}.self_attn.{name}.weight"]
        y = _mm(x, w, self.dtype)
        if self.use_lora and name in ("q_proj", "v_proj"):
            p = f"base_model.model.model.layers.{layer}.self_attn.{name}."
            a = _mm(x, self.sd[p + "lora_A.weight"], self.dtype)
            b = _mm(a, self.sd[p + "lora_B.weight"], self.dtype)
            y = y + b * self.cfg.lora_scaling
        return y

    def embed(self, input_ids: torch.Tensor, img_embeds: Optional[torch.Tensor]) -> torch.Tensor:
        """LlamaModel.forward splice, modeling_llama_imgemb.py:571-594 + split_at_img :498-520.
        img_embeds: [B,32,768] (Q-Former output, any float dtype) or None (plain embedding lookup)."""
        E = self.sd["model.embed_tokens.weight"]
        if img_embeds is None:
            return E[input_ids]
        w, b = self.sd["model.img_proj_layer.weight"], self.sd["model.img_proj_layer.bias"]
        img = (img_embeds.to(self.dtype).float() @ w.float().t() + b.float()).to(self.dtype)   # :577/:579
        rows, cols = (input_ids == IMG_TOKEN_ID).nonzero(as_tuple=True)
        rows, cols = rows[::NUM_IMG_TOKENS], cols[::NUM_IMG_TOKENS]
        pos = torch.zeros(input_ids.size(0), dtype=torch.long)
        pos[rows] = cols                                       # rows without <IMG> default to 0 (:507-510)
        out = []
        for i in range(input_ids.size(0)):
            p = int(pos[i])
            out.append(torch.cat([E[input_ids[i, :p]], img[i], E[input_ids[i, p + NUM_IMG_TOKENS:]]], dim=0))
        return torch.stack(out, 0)

    def layer(self, x, li: int, mask, position_ids, past: Optional[Tuple[torch.Tensor, torch.Tensor]]):
        """LlamaDecoderLayer.forward :266-318 (residual adds in the model dtype)."""
        cfg, dt = self.cfg, self.dtype
        B, q_len, H = x.shape
        nh, hd = cfg.num_attention_heads, cfg.head_dim
        p = f"model.layers.{li}."
        h = rmsnorm(x, self.sd[p + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = self._proj(h, li, "q_proj").view(B, q_len, nh, hd).transpose(1, 2)
        k = self._proj(h, li, "k_proj").view(B, q_len, nh, hd).transpose(1, 2)
        v = self._proj(h, li, "v_proj").view(B, q_len, nh, hd).transpose(1, 2)
        q, k = apply_rope(q, k, self.cos, self.sin, position_ids)
        if past is not None:                                   # :209-212 cache holds post-RoPE K
            k = torch.cat([past[0], k], dim=2)
            v = torch.cat([past[1], v], dim=2)
        a = attention(q, k, v, mask, dt).transpose(1, 2).reshape(B, q_len, H)
        x = x + _mm(a, self.sd[p + "self_attn.o_proj.weight"], dt)
        h = rmsnorm(x, self.sd[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        g = _mm(h, self.sd[p + "mlp.gate_proj.weight"], dt)
        u = _mm(h, self.sd[p + "mlp.up_proj.weight"], dt)
        x = x + _mm(F.silu(g) * u, self.sd[p + "mlp.down_proj.weight"], dt)     # :158-159
        return x, (k, v)

    def forward(self, input_ids, attention_mask, position_ids, past=None, img_embeds=None,
                return_hidden: bool = False):
        """LlamaForCausalLM.forward :705-793 -> logits [B,q,V] (all positions, like the reference)."""
        x = self.embed(input_ids, img_embeds if past is None else None)
        q_len = input_ids.shape[1]
        mask = make_attention_mask(attention_mask, q_len, self.dtype)
        new_past, hiddens = [], []
        for li in range(self.cfg.num_hidden_layers):
            if return_hidden:
                hiddens.append(x)
            x, kv = self.layer(x, li, mask, position_ids, None if past is None else past[li])
            new_past.append(kv)
        x = rmsnorm(x, self.sd["model.norm.weight"], self.cfg.rms_norm_eps)
        logits = _mm(x, self.sd["lm_head.weight"], self.dtype)
        if return_hidden:
            hiddens.append(x)
            return logits, new_past, hiddens
        return logits, new_past

    # -- generation ---------------------------------------------------------------------
    @staticmethod
    def positions_from_mask(attention_mask: torch.Tensor) -> torch.Tensor:
        """prepare_inputs_for_generation :804-808: cumsum(mask)-1, pads forced to 1."""
        pos = attention_mask.long().cumsum(-1) - 1
        pos.masked_fill_(attention_mask == 0, 1)
        return pos

    def generate(self, input_ids: torch.Tensor, img_embeds: Optional[torch.Tensor], max_new_tokens: int,
                 suppress_eos: bool = False, return_scores: bool = False):
        """HF transformers==4.28.1 ``GenerationMixin.greedy_search`` restated over
        prepare_inputs_for_generation (:795-836): attention mask inferred as ``ids != pad`` (pad=0), fp16
        argmax without upcast, finished rows emit pad, stop when all rows hit EOS or the length cap."""
        cfg = self.cfg
        ids = input_ids.clone()
        mask = ids.ne(cfg.pad_token_id).long()
        unfinished = torch.ones(ids.shape[0], dtype=torch.long)
        past, scores = None, []
        for _ in range(max_new_tokens):
            pos = self.positions_from_mask(mask)
            if past is None:
                logits, past = self.forward(ids, mask, pos, None, img_embeds)
            else:
                logits, past = self.forward(ids[:, -1:], mask, pos[:, -1:], past, None)
            nxt_logits = logits[:, -1, :]
            if return_scores:
                scores.append(nxt_logits)
            if suppress_eos:
                nxt_logits = nxt_logits.clone()
                nxt_logits[:, cfg.eos_token_id] = torch.finfo(nxt_logits.dtype).min
            tok = torch.argmax(nxt_logits, dim=-1)
            tok = tok * unfinished + cfg.pad_token_id * (1 - unfinished)
            ids = torch.cat([ids, tok[:, None]], dim=-1)
            mask = torch.cat([mask, mask.new_ones((mask.shape[0], 1))], dim=-1)
            unfinished = unfinished.mul((tok != cfg.eos_token_id).long())
            if unfinished.max() == 0:
                break
        return (ids, scores) if return_scores else ids


def llama_beam_search(orc: "LlamaOracle", input_ids: torch.Tensor, img_embeds: Optional[torch.Tensor], max_new_tokens: int,
                      num_beams: int, length_penalty: float = 1.0, early_stopping=False, length_norm: str = "full"):
    """generate(num_beams=k) over the oracle model: the cache rows follow ``_reorder_cache`` (index_select on dim 0).
    Image rows are expanded with the prompts (the reference's own ``dicom`` list is NOT expanded by HF and its forward then
    fails on the shape mismatch - SURVEY.md 8f row 4; text-only prompts are the reference's working case)."""
    cfg = orc.cfg
    state = {"past": None, "mask": None}
    img = None if img_embeds is None else img_embeds.repeat_interleave(num_beams, dim=0)

    def step(ids, beam_idx):
        if state["past"] is None:
            mask = ids.ne(cfg.pad_token_id).long()
            logits, past = orc.forward(ids, mask, orc.positions_from_mask(mask), None, img)
        else:
            past = [(k.index_select(0, beam_idx), v.index_select(0, beam_idx)) for k, v in state["past"]]
            mask = torch.cat([state["mask"], state["mask"].new_ones((ids.shape[0], 1))], dim=-1)
            pos = orc.positions_from_mask(mask)
            logits, past = orc.forward(ids[:, -1:], mask, pos[:, -1:], past, None)
        state["past"], state["mask"] = past, mask
        return logits[:, -1, :]

    return beam_search(step, input_ids, num_beams, input_ids.shape[1] + max_new_tokens, cfg.pad_token_id, cfg.eos_token_id,
                       length_penalty, early_stopping, length_norm)


# ======================================================================================
# Beam search (transformers==4.28.1 GenerationMixin.beam_search + BeamSearchScorer, third-party, restated;
# reached through generate(num_beams=...) at test.py:467,629; cache reordering = _reorder_cache,
# modeling_llama_imgemb.py:838-843)
# ======================================================================================

class _BeamHypotheses:
    """transformers 4.28.1 generation/beam_search.py BeamHypotheses (n-best list of finished hypotheses of one batch item)."""

    def __init__(self, num_beams: int, length_penalty: float, early_stopping, max_length: Optional[int], prompt_len: int = 0):
        self.length_penalty, self.early_stopping, self.max_length, self.num_beams = length_penalty, early_stopping, max_length, num_beams
        self.beams: List[Tuple[float, torch.Tensor]] = []
        self.worst_score = 1e9
        # 4.28.1 (the reference's pin) normalises by the FULL hypothesis length, prompt included; later transformers releases
        # normalise by the generated length only.  prompt_len > 0 selects the later rule - used ONLY to check this restatement
        # against the transformers build installed here (oracle/make_golden_beam.py); 0 = the reference's behaviour.
        self.prompt_len = prompt_len

    def add(self, hyp: torch.Tensor, sum_logprobs: float, eos_counts: bool = False):
        # (the later rule also counts the EOS token of a finished hypothesis; 4.28.1 stores and measures the hypothesis without it)
        n = hyp.shape[-1] - self.prompt_len + (1 if (eos_counts and self.prompt_len) else 0)
        score = sum_logprobs / (n ** self.length_penalty)
        if len(self.beams) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self.beams) > self.num_beams:
                order = sorted([(sc, idx) for idx, (sc, _) in enumerate(self.beams)])
                del self.beams[order[0][1]]
                self.worst_score = order[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        if len(self.beams) < self.num_beams:
            return False
        if self.early_stopping is True:
            return True
        if self.early_stopping is False:
            return self.worst_score >= best_sum_logprobs / (cur_len - self.prompt_len) ** self.length_penalty
        if self.length_penalty > 0.0:            # "never"
            return self.worst_score >= best_sum_logprobs / (self.max_length - self.prompt_len) ** self.length_penalty
        return self.worst_score >= best_sum_logprobs / (cur_len - self.prompt_len) ** self.length_penalty


def beam_search(step_logits, input_ids: torch.Tensor, num_beams: int, max_length: int, pad_token_id: int, eos_token_id: int,
                length_penalty: float = 1.0, early_stopping=False, length_norm: str = "full"):
    """GenerationMixin.beam_search of transformers 4.28.1 (num_return_sequences = 1, no logits processors), restated.

    ``step_logits(input_ids [B*nb, L], beam_idx or None) -> logits [B*nb, V]`` runs the model on the last position (the whole
    prompt on the first call) after reordering its cache rows by ``beam_idx`` (``_reorder_cache``).  ``input_ids`` is the
    UN-expanded prompt [B, T].  Returns (sequences [B, <= max_length], sequence_scores [B])."""
    B = input_ids.shape[0]
    ids = input_ids.repeat_interleave(num_beams, dim=0)                       # _expand_inputs_for_generation
    assert length_norm in ("full", "generated")
    plen = input_ids.shape[1] if length_norm == "generated" else 0
    hyps = [_BeamHypotheses(num_beams, length_penalty, early_stopping, max_length, plen) for _ in range(B)]
    done = [False] * B
    beam_scores = torch.zeros((B, num_beams), dtype=torch.float)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    beam_idx = None
    while True:
        logits = step_logits(ids, beam_idx)
        scores = F.log_softmax(logits, dim=-1)                                 # in the logits dtype, like HF
        scores = scores.cpu() + beam_scores[:, None]                           # fp32 by type promotion
        V = scores.shape[-1]
        # torch.topk leaves the order of EQUAL scores unspecified (and fp16 log-probs tie often); a stable descending sort fixes
        # it to "lowest flat index first" - one valid instance of the transformers behaviour, and reproducible across devices
        srt_scores, srt_tokens = torch.sort(scores.view(B, num_beams * V), dim=1, descending=True, stable=True)
        top_scores, top_tokens = srt_scores[:, : 2 * num_beams], srt_tokens[:, : 2 * num_beams]
        top_idx = torch.div(top_tokens, V, rounding_mode="floor")
        top_tokens = top_tokens % V
        # ---- BeamSearchScorer.process -------------------------------------------------------------------------------
        cur_len = ids.shape[-1]
        nxt_scores = torch.zeros((B, num_beams), dtype=top_scores.dtype)
        nxt_tokens = torch.zeros((B, num_beams), dtype=top_tokens.dtype)
        nxt_idx = torch.zeros((B, num_beams), dtype=top_idx.dtype)
        ids_cpu = ids.cpu()
        for b in range(B):
            if done[b]:
                nxt_scores[b, :], nxt_tokens[b, :], nxt_idx[b, :] = 0, pad_token_id, 0
                continue
            k = 0
            for rank in range(2 * num_beams):
                tok, sc, idx = int(top_tokens[b, rank]), top_scores[b, rank], int(top_idx[b, rank])
                row = b * num_beams + idx
                if tok == eos_token_id:
                    if rank >= num_beams:
                        continue
                    hyps[b].add(ids_cpu[row].clone(), float(sc), eos_counts=True)
                else:
                    nxt_scores[b, k], nxt_tokens[b, k], nxt_idx[b, k] = sc, tok, row
                    k += 1
                if k == num_beams:
                    break
            if k < num_beams:
                raise ValueError(f"At most {num_beams} tokens in the top {2 * num_beams} can be equal to `eos_token_id`")
            # 4.28.1 passes the length BEFORE this step's token is appended; the generated-length rule of later releases counts it
            done[b] = done[b] or hyps[b].is_done(float(top_scores[b].max()), cur_len + (1 if plen else 0))
        beam_scores, beam_tokens, beam_idx = nxt_scores.view(-1), nxt_tokens.view(-1), nxt_idx.view(-1)
        ids = torch.cat([ids_cpu[beam_idx, :], beam_tokens[:, None]], dim=-1).to(ids.device)
        if all(done) or ids.shape[-1] >= max_length:
            break
    # ---- BeamSearchScorer.finalize -----------------------------------------------------------------------------------
    ids_cpu = ids.cpu()
    for b in range(B):
        if done[b]:
            continue
        for k in range(num_beams):
            hyps[b].add(ids_cpu[b * num_beams + k], float(beam_scores[b * num_beams + k]))
    best, best_scores = [], torch.zeros(B, dtype=torch.float32)
    for b in range(B):
        sc, hyp = sorted(hyps[b].beams, key=lambda x: x[0]).pop()
        best.append(hyp)
        best_scores[b] = sc
    lens = [len(x) for x in best]
    sent_max_len = min(max(lens) + 1, max_length)
    decoded = torch.full((B, sent_max_len), pad_token_id, dtype=ids_cpu.dtype) if min(lens) != max(lens) else \
        torch.zeros((B, sent_max_len), dtype=ids_cpu.dtype)
    for b, hyp in enumerate(best):
        decoded[b, : lens[b]] = hyp
        if lens[b] < sent_max_len:
            decoded[b, lens[b]] = eos_token_id
    return decoded, best_scores


# ======================================================================================
# Vision trunk + projector (biovil_t/*, torchvision ResNet-50 v1.5)
# ======================================================================================

def _bn(x, sd, name, eps):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=eps)


def resnet_trunk(x: torch.Tensor, sd, cfg, prefix="visual_encoder.encoder.encoder.") -> torch.Tensor:
    """ResNetHIML.forward biovil_t/resnet.py:25-47 over torchvision ``Bottleneck`` (v1.5: stride on the 3x3):
    conv1 7x7/2 -> BN -> ReLU -> maxpool 3x3/2 -> layer1..4.  Eval-mode BN (running stats)."""
    x = F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn(x, sd, prefix + "bn1", cfg.bn_eps))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nblocks in enumerate(cfg.layers):
        for b in range(nblocks):
            p = f"{prefix}layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            idt = x
            o = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1", cfg.bn_eps))
            o = F.relu(_bn(F.conv2d(o, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2", cfg.bn_eps))
            o = _bn(F.conv2d(o, sd[p + ".conv3.weight"]), sd, p + ".bn3", cfg.bn_eps)
            if (p + ".downsample.0.weight") in sd:
                idt = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1", cfg.bn_eps)
            x = F.relu(o + idt)
    return x


def sine_position_embedding(H: int, W: int, embedding_dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """SinePositionEmbedding(embedding_dim, normalize=True)(mask=ones[1,H,W]) biovil_t/transformer.py:225-266 -> [1, H*W, 2*dim]."""
    mask = torch.ones(1, H, W)
    y_embed, x_embed = mask.cumsum(1, dtype=torch.float32), mask.cumsum(2, dtype=torch.float32)
    scale = 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + 1e-6) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + 1e-6) * scale
    dim_t = torch.arange(embedding_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / embedding_dim)
    pos_x, pos_y = x_embed[:, :, :, None] / dim_t, y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).view(1, H * W, embedding_dim * 2)


def vit_pooler(cur: torch.Tensor, prev: torch.Tensor, sd, cfg, prefix="visual_encoder.encoder.vit_pooler.") -> torch.Tensor:
    """VisionTransformerPooler.forward / forward_after_reshape (biovil_t/transformer.py:77-118) in eval mode: tokens of the
    current and the previous image concatenated, sine position + type embeddings added to the NORMALISED input of every block's
    attention (Block.forward :213-218: q = k = v = norm1(x) + emb), pre-LN blocks with MultiHeadAttentionLayer (:148-166, scale
    after QK^T) and timm Mlp (exact GELU); returns the current image's tokens after norm_post, back in [B,C,H,W]."""
    B, C, H, W = cur.shape
    L = H * W
    x = torch.cat([cur.view(B, C, L).transpose(1, 2), prev.view(B, C, L).transpose(1, 2)], dim=1)          # [B, 2L, C]
    pos = sine_position_embedding(H, W, C // 2).repeat(B, 1, 1)
    te = sd[prefix + "type_embed"]
    emb = torch.cat([pos, pos], dim=1) + torch.cat([te[0].expand(B, L, -1), te[1].expand(B, L, -1)], dim=1)
    nh = cfg.pooler_heads
    hd = C // nh
    for i in range(cfg.pooler_blocks):
        p = prefix + f"blocks.{i}."
        xe = _ln(x, sd, p + "norm1", cfg.pooler_ln_eps) + emb

        def heads(t):
            return t.reshape(B, 2 * L, nh, hd).permute(0, 2, 1, 3)

        q = heads(F.linear(xe, sd[p + "attn.proj_q.weight"]))
        k = heads(F.linear(xe, sd[p + "attn.proj_k.weight"]))
        v = heads(F.linear(xe, sd[p + "attn.proj_v.weight"]))
        attn = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, 2 * L, C)
        x = x + _lin(o, sd, p + "attn.proj")
        h = _ln(x, sd, p + "norm2", cfg.pooler_ln_eps)
        x = x + _lin(F.gelu(_lin(h, sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    x = _ln(x, sd, prefix + "norm_post", cfg.pooler_ln_eps)
    return x[:, :L].transpose(1, 2).reshape(B, C, H, W)


def image_model(x: torch.Tensor, sd, cfg, previous: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ImageModel.forward biovil_t/model.py:76-91 over MultiImageEncoder.forward biovil_t/encoder.py:110-136: trunk ->
    backbone_to_vit 1x1 -> concat with either the broadcast missing_previous_emb (single image, :124-130) or the
    VisionTransformerPooler output over (current, previous) (:117-123) -> projector MLP (modules.py:43-47).
    Returns projected_patch_embeddings [B,J,g,g] (NCHW)."""
    E = "visual_encoder.encoder."
    P = "visual_encoder.projector.model."
    B = x.shape[0]
    if previous is not None:
        assert previous.shape == x.shape
        both = F.conv2d(resnet_trunk(torch.cat([x, previous], dim=0), sd, cfg), sd[E + "backbone_to_vit.weight"])
        patch = both[:B]
        diff = vit_pooler(patch, both[B:], sd, cfg)
    else:
        patch = F.conv2d(resnet_trunk(x, sd, cfg), sd[E + "backbone_to_vit.weight"])
        _, _, W, Hh = patch.shape
        diff = sd[E + "missing_previous_emb"].repeat(B, 1, W, Hh)
    fused = torch.cat([patch, diff], dim=1)
    h = F.relu(_bn(F.conv2d(fused, sd[P + "0.weight"]), sd, P + "1", cfg.bn_eps))
    return F.conv2d(h, sd[P + "3.weight"], sd[P + "3.bias"])


# ======================================================================================
# Q-Former (Qformer.py query-only branch) and forward_image glue (blip2_qformer.py:467-484)
# ======================================================================================

def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _bert_attention(sd, prefix, hidden, kv_src, n_heads, eps):
    """BertAttention = BertSelfAttention.forward Qformer.py:169-275 (+0 masks, scale after QK^T :244) +
    BertSelfOutput :285-289 (dense + residual + post-LN)."""
    B, Lq, Hq = hidden.shape
    hd = Hq // n_heads

    def split(t):
        return t.view(B, -1, n_heads, hd).permute(0, 2, 1, 3)

    k = split(_lin(kv_src, sd, prefix + "self.key"))
    v = split(_lin(kv_src, sd, prefix + "self.value"))
    q = split(_lin(hidden, sd, prefix + "self.query"))
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
    p = torch.softmax(s, dim=-1)
    ctx = torch.matmul(p, v).permute(0, 2, 1, 3).contiguous().view(B, Lq, Hq)
    return _ln(_lin(ctx, sd, prefix + "output.dense") + hidden, sd, prefix + "output.LayerNorm", eps)


def qformer(image_embeds: torch.Tensor, sd, cfg) -> torch.Tensor:
    """BertModel.forward Qformer.py:804-965 with input_ids=None: embeddings = LayerNorm(query_tokens) (:78-108),
    all masks zero, 12x BertLayer.forward :402-484 with query_length=32 (self-attn, cross-attn on even layers,
    intermediate_query/output_query FFN with exact-erf GELU)."""
    Bp = "Qformer.bert."
    B = image_embeds.shape[0]
    h = _ln(sd["query_tokens"].expand(B, -1, -1), sd, Bp + "embeddings.LayerNorm", cfg.q_ln_eps)
    for i in range(cfg.q_layers):
        p = Bp + f"encoder.layer.{i}."
        h = _bert_attention(sd, p + "attention.", h, h, cfg.q_heads, cfg.q_ln_eps)
        if i % cfg.cross_attention_freq == 0:
            h = _bert_attention(sd, p + "crossattention.", h, image_embeds, cfg.q_heads, cfg.q_ln_eps)
        inter = F.gelu(_lin(h, sd, p + "intermediate_query.dense"))
        h = _ln(_lin(inter, sd, p + "output_query.dense") + h, sd, p + "output_query.LayerNorm", cfg.q_ln_eps)
    return h


def forward_image(image: torch.Tensor, sd, cfg, previous_image: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Blip2Qformer.forward_image blip2_qformer.py:467-484 (fp32): the ``reshape(B,-1,1408)`` at :469 is a raw
    reinterpretation of the NCHW buffer, not a permute; ``ln_vision`` is blip2.py:199-205.  ``previous_image`` selects the
    two-image branch of the BioViL-T encoder (the reference's forward_image never passes one: SURVEY.md 8f row 4)."""
    proj = image_model(image.float(), sd, cfg, None if previous_image is None else previous_image.float())
    image_embeds = proj.reshape(image.shape[0], -1, cfg.joint_feature_size)
    image_embeds = _ln(image_embeds, sd, "ln_vision", cfg.ln_vision_eps)
    return qformer(image_embeds, sd, cfg), image_embeds


# ======================================================================================
# Whole path
# ======================================================================================

def image_to_report(images, prompts, vis_sd, vis_cfg, llm: LlamaOracle, max_new_tokens: int,
                    suppress_eos: bool = False):
    """demo.py:269-297 / test.py:336-348 in one call: forward_image -> splice -> greedy decode."""
    q_out, _ = forward_image(images, vis_sd, vis_cfg)
    return llm.generate(prompts, q_out, max_new_tokens, suppress_eos=suppress_eos)


class Prompter:
    """utils/prompter.py:10-50 restated with the template passed in (the reference reads
    data/templates/<name>.json relative to CWD)."""

    def __init__(self, template: Dict[str, str]):
        self.template = template

    def generate_prompt(self, instruction, input=None, label=None):
        res = (self.template["prompt_input"].format(instruction=instruction, input=input) if input
               else self.template["prompt_no_input"].format(instruction=instruction))
        return f"{res}{label}" if label else res

    def get_response(self, output: str) -> str:
        return output.split(self.template["response_split"])[-1].strip()
