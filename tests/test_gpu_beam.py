"""Beam search through the product path (generate(num_beams=k), test.py:467,629) against the oracle's restatement of
transformers 4.28.1 beam_search (pinned against transformers' own beam search in tests/test_oracle_golden.py).

The engine side under test: prefill over the k-times expanded prompts, rd_llm_reorder_cache (_reorder_cache,
modeling_llama_imgemb.py:838-843), rd_llm_force_tokens, eager decode steps.  fp16 near-ties can reorder candidates whose
accumulated scores differ by less than the logit noise, so sequences are required to be equal only where the oracle's final
best and second-best hypotheses are separated by more than that noise; the best score must agree to the noise in any case."""
import pytest
import torch

from radialog_b200 import synth
from radialog_b200.llm import LlamaForCausalLM
from oracle import radialog_oracle as O

pytestmark = pytest.mark.gpu


def build(cfg, dtype, dev, seed, eos_boost=1.0, lora=True):
    sd = synth.make_llama_weights(cfg, seed=seed, dtype=torch.float32, lora=lora)
    sd = {k: v.to(torch.float16).float() for k, v in sd.items()}
    sd["lm_head.weight"][cfg.eos_token_id] *= eos_boost
    return LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev), O.LlamaOracle(cfg, sd, dtype)


def oracle_sequence_scores(orc, prompts, img, seqs, eos):
    """Length-normalised log-probability (transformers 4.28.1 rule: sum of token log-probs incl. a closing EOS, divided by the
    hypothesis length without it, prompt included) that the ORACLE model assigns to each returned sequence."""
    T = prompts.shape[1]
    out = []
    for b in range(seqs.shape[0]):
        gen = seqs[b, T:].tolist()
        n = gen.index(eos) + 1 if eos in gen else len(gen)
        full = seqs[b:b + 1, :T + n]
        mask = full.ne(0).long()
        mask[:, T:] = 1
        logits, _ = orc.forward(full[:, :-1], mask[:, :-1], orc.positions_from_mask(mask[:, :-1]), None, None if img is None else img[b:b + 1])
        lp = torch.log_softmax(logits[0, T - 1:, :], dim=-1).float()            # in the logits dtype like generate, then accumulated in fp32
        tok = full[0, T:]
        total = lp.gather(1, tok[:, None]).sum().item()
        out.append(total / (T + n - (1 if eos in gen else 0)))
    return torch.tensor(out)


@pytest.mark.parametrize("nb,eos_boost,with_img", [(3, 1.0, False), (4, 4.0, False), (2, 1.0, True), (3, 4.0, True)])
def test_beam_search_matches_oracle(cuda_dev, nb, eos_boost, with_img):
    dtype = torch.float16
    cfg = synth.tiny_llama_cfg()
    model, orc = build(cfg, dtype, cuda_dev, seed=3, eos_boost=eos_boost)
    B, new = 3, 14
    prompts = synth.make_prompts(B, seed=50 + nb, ragged=True)
    img = None
    if with_img:
        g = torch.Generator().manual_seed(9)
        img = (torch.randn(B, 32, cfg.qformer_hidden, generator=g) * 0.5).float()
    else:
        prompts = torch.where(prompts == synth.IMG_TOKEN_ID, torch.full_like(prompts, 77), prompts)      # text-only prompt
    o_seq, o_sc = O.llama_beam_search(orc, prompts, img, new, nb)
    out = model.generate(prompts.to(cuda_dev), img_embeds=None if img is None else img.to(cuda_dev), max_new_tokens=new, num_beams=nb,
                         return_dict_in_generate=True, output_scores=True)
    seq, sc = out.sequences.cpu(), out.sequences_scores.cpu()
    assert seq.shape[0] == B and torch.equal(seq[:, :prompts.shape[1]], prompts)
    assert len(out.scores) >= 1 and out.scores[0].shape == (B * nb, cfg.vocab_size)
    # accumulated log-prob noise: ~1e-2 per step in fp16 at this logit scale, normalised by the full length
    tol = 1e-2 * new / prompts.shape[1]
    # Beam search prunes on near-ties, so fp16 noise can steer it to a different - better or worse - final hypothesis.  Per row:
    # either the oracle's hypothesis exactly, or a hypothesis (a) whose reported score is what the ORACLE model assigns to that
    # very sequence (score bookkeeping, cache reordering and token hand-over are right) and (b) that is as good as the oracle's
    # best up to a few noise quanta.
    re_scored = oracle_sequence_scores(orc, prompts, img, seq, cfg.eos_token_id)
    exact = 0
    for b in range(B):
        assert abs(re_scored[b] - sc[b]) <= tol, f"row {b}: reported score {sc[b]:.5f} vs oracle re-score of the same sequence {re_scored[b]:.5f}"
        if seq.shape == o_seq.shape and torch.equal(seq[b], o_seq[b]):
            exact += 1
            assert abs(sc[b] - o_sc[b]) <= tol
        else:
            assert re_scored[b] >= o_sc[b] - 4 * tol, (f"row {b}: returned hypothesis scores {re_scored[b]:.5f} under the oracle model, "
                                                        f"the oracle's best {o_sc[b]:.5f}:\n{seq[b]}\n{o_seq[b]}")
    assert exact >= 1, f"no row reproduced the oracle's hypothesis:\n{seq}\n{o_seq}"
    # a greedy call afterwards must still work (captured graphs were dropped, the cache buffers were swapped)
    greedy = model.generate(prompts.to(cuda_dev), img_embeds=None if img is None else img.to(cuda_dev), max_new_tokens=6, suppress_eos=True)
    o_greedy = orc.generate(prompts, img, 6, suppress_eos=True)
    assert (greedy.cpu() == o_greedy).float().mean().item() > 0.9


def test_beam_search_k1_is_greedy_and_bad_args(cuda_dev):
    cfg = synth.tiny_llama_cfg()
    model, orc = build(cfg, torch.float16, cuda_dev, seed=5)
    prompts = synth.make_prompts(2, seed=1)
    img = torch.zeros(2, 32, cfg.qformer_hidden)
    with pytest.raises(ValueError):
        model.generate(prompts.to(cuda_dev), img_embeds=img.to(cuda_dev), max_new_tokens=4, num_beams=2, reuse_cache=True)
