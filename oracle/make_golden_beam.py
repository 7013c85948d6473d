#!/usr/bin/env python
"""Golden vectors for beam search: sequences / sequence scores of transformers' OWN beam search.

The reference reaches beam search through ``lang_model.generate(..., num_beams=k)`` (test.py:467,629), i.e. transformers
4.28.1 ``GenerationMixin.beam_search`` + ``BeamSearchScorer`` - third-party code absent from /root/reference, restated in
oracle/radialog_oracle.py.  To pin that restatement this script runs the beam search of the transformers build installed in
THIS container (stock ``LlamaForCausalLM``, eager attention, fp32, CPU) on seeded tiny LLaMA weights under the reference's key
names (text-only prompts: the reference's working beam case) and stores prompts + outputs in tests/golden/beam_tiny_f32.npz.
    python -m oracle.make_golden_beam
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radialog_b200 import synth  # noqa: E402


def main():
    import transformers
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = synth.tiny_llama_cfg()
    seed, eos_boost = 3, 4.0
    sd = synth.make_llama_weights(cfg, seed=seed, dtype=torch.float32, lora=False)
    sd = {k: v for k, v in sd.items() if "img_proj" not in k}
    sd["lm_head.weight"][cfg.eos_token_id] *= eos_boost            # EOS reachable: finished hypotheses, padding, early finish
    hc = LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_attention_heads, max_position_embeddings=cfg.max_position_embeddings,
                     rms_norm_eps=cfg.rms_norm_eps, pad_token_id=0, bos_token_id=1, eos_token_id=2, rope_theta=10000.0,
                     attention_bias=False, tie_word_embeddings=False, attn_implementation="eager")
    model = LlamaForCausalLM(hc).eval()
    missing = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if "rotary" not in k], missing
    g = torch.Generator().manual_seed(11)
    out = {}
    for name, B, T, nb, new in (("a", 3, 12, 3, 16), ("b", 2, 20, 4, 24)):
        ids = torch.randint(3, 32000, (B, T), generator=g)
        ids[0, :3] = 0                                                 # left padding on one row
        with torch.no_grad():
            r = model.generate(input_ids=ids, attention_mask=ids.ne(0).long(), num_beams=nb, max_new_tokens=new, do_sample=False,
                               length_penalty=1.0, early_stopping=False, return_dict_in_generate=True, output_scores=True,
                               pad_token_id=0, eos_token_id=2)
        out[f"{name}_prompts"] = ids.numpy()
        out[f"{name}_sequences"] = r.sequences.numpy()
        out[f"{name}_scores"] = r.sequences_scores.numpy()
        out[f"{name}_cfg"] = np.array([nb, new])
    out["seed"], out["eos_boost"] = np.array(seed), np.array(eos_boost)
    out["transformers_version"] = np.array(transformers.__version__)
    path = os.path.join(ROOT, "tests", "golden", "beam_tiny_f32.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
