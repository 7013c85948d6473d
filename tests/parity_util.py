"""Shared parity helpers of the GPU tests.

Two harnesses:

* ``assert_ids_match`` - free-running greedy ids against the oracle's (small models, CPU oracle).
* ``teacher_forced_parity`` - SURVEY.md section 7 "hard parts": at every step BOTH implementations consume the ORACLE's
  token, so one near-tie cannot hide everything after it.  Every step of every row is checked: ``max|dlogit|`` against the
  stated tolerance or twice the MEASURED noise floor of the reference arithmetic itself (whichever is larger), and
  ``argmax(engine) == argmax(oracle)`` whenever the oracle's top-2 margin exceeds twice the measured ``|dlogit|`` of that
  row (both candidates can move by that much); steps below that margin are counted and reported, never skipped silently.

``CudaOracle`` runs ``oracle/radialog_oracle.py`` unchanged on the GPU (so a 32-layer, 128-token run takes seconds): same
rounding points, matmuls as fp32 SGEMM of the exactly-representable fp16/bf16 operands (TF32 off) = "fp32 accumulate, one
rounding", i.e. what the CPU oracle computes up to fp32 summation order.  ``check_cuda_oracle_against_cpu`` pins that claim.
"""
import contextlib

import torch

from oracle import radialog_oracle as O


def assert_ids_match(ids, ref_ids, ref_scores, prompt_len, dtype, what, min_exact_rows=0.5):
    """ids equal to ref_ids; a row may diverge only at a step where the reference's top-2 margin is a near tie."""
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    B = ids.shape[0]
    n = min(ids.shape[1], ref_ids.shape[1]) - prompt_len
    exact_rows = 0
    for b in range(B):
        row_ok = True
        for s in range(n):
            if ids[b, prompt_len + s] != ref_ids[b, prompt_len + s]:
                top = ref_scores[s][b].float().topk(2).values
                margin = (top[0] - top[1]).item()
                tol = 3 * ulp * max(1.0, top[0].abs().item())
                assert margin <= tol, (f"{what}: row {b} step {s}: got {ids[b, prompt_len + s].item()} expected "
                                       f"{ref_ids[b, prompt_len + s].item()} with margin {margin:.4g} > tie tolerance {tol:.4g}")
                row_ok = False
                break
        exact_rows += row_ok
    assert exact_rows >= min_exact_rows * B, f"{what}: only {exact_rows}/{B} rows bit-equal"


class CudaOracle:
    """oracle.LlamaOracle with its tensors on the GPU and ``_mm`` as an fp32 SGEMM over cached fp32 copies of the weights."""

    def __init__(self, cfg, sd, dtype, device, cache_fp32=True):
        assert not torch.backends.cuda.matmul.allow_tf32, "the oracle's fp32 matmuls must not run in TF32"
        self.device = torch.device(device)
        self.dtype = dtype
        self.orc = O.LlamaOracle(cfg, {k: v.to(self.device) for k, v in sd.items()}, dtype)
        self.orc.cos, self.orc.sin = self.orc.cos.to(self.device), self.orc.sin.to(self.device)
        self._w32 = {} if cache_fp32 else None

    @contextlib.contextmanager
    def _ctx(self):
        old = O._mm
        cache = self._w32

        def mm(x, w, dtype):
            if cache is None:
                wf = w.float()
            else:
                wf = cache.get(w.data_ptr())
                if wf is None:
                    wf = cache[w.data_ptr()] = w.float()
            return (x.float() @ wf.t()).to(dtype)

        O._mm = mm
        try:
            with torch.device(self.device), torch.no_grad():
                yield
        finally:
            O._mm = old

    def generate(self, prompts, img, n_new, suppress_eos=True):
        with self._ctx():
            ids, scores = self.orc.generate(prompts.to(self.device), None if img is None else img.to(self.device), n_new,
                                            suppress_eos=suppress_eos, return_scores=True)
        return ids, scores

    def forward(self, ids, img=None):
        with self._ctx():
            ids = ids.to(self.device)
            mask = ids.ne(0).long()
            logits, _ = self.orc.forward(ids, mask, self.orc.positions_from_mask(mask), None, None if img is None else img.to(self.device))
        return logits

    def release(self):
        self._w32 = {} if self._w32 is not None else None


def check_cuda_oracle_against_cpu(cfg, sd, dtype, device, prompts, img, n_new=6):
    """The GPU run of the oracle against its CPU run (the pinned one): same ids, logits equal to fp32 summation-order noise."""
    cpu = O.LlamaOracle(cfg, sd, dtype)
    c_ids, c_scores = cpu.generate(prompts, img, n_new, suppress_eos=True, return_scores=True)
    g = CudaOracle(cfg, sd, dtype, device)
    g_ids, g_scores = g.generate(prompts, img, n_new)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    worst = 0.0
    for s in range(n_new):
        a, b = c_scores[s].float(), g_scores[s].float().cpu()
        worst = max(worst, ((a - b).abs().max() / a.abs().max()).item())
        if not torch.equal(c_ids[:, :prompts.shape[1] + s + 1], g_ids.cpu()[:, :prompts.shape[1] + s + 1]):
            break
    assert worst <= 4 * ulp, f"oracle on cuda differs from the oracle on the CPU by {worst:.3e} relative"
    return worst


def oracle_forced_scores(orc: CudaOracle, prompts, img, forced, backend):
    """The oracle along a FORCED token trajectory (same loop as LlamaOracle.generate, next token = forced[:, s]) with its
    matmuls on `backend`: "sgemm" (fp32 SGEMM, the checker) or "cublas" (model-dtype tensor-core GEMM, what the reference's
    nn.Linear runs on a GPU).  The gap between the two is the noise floor of the reference arithmetic itself."""
    import torch.nn.functional as F
    dev = orc.device
    n = forced.shape[1]
    scores = []
    with orc._ctx():
        if backend == "cublas":
            O._mm = lambda x, w, dt: F.linear(x, w)
        o = orc.orc
        ids = prompts.to(dev)
        mask = ids.ne(o.cfg.pad_token_id).long()
        past = None
        for s in range(n):
            pos = o.positions_from_mask(mask)
            if past is None:
                logits, past = o.forward(ids, mask, pos, None, None if img is None else img.to(dev))
            else:
                logits, past = o.forward(ids[:, -1:], mask, pos[:, -1:], past, None)
            scores.append(logits[:, -1, :])
            ids = torch.cat([ids, forced[:, s:s + 1]], dim=-1)
            mask = torch.cat([mask, mask.new_ones((mask.shape[0], 1))], dim=-1)
    return scores


def teacher_forced_parity(model, orc: CudaOracle, prompts, img, n_new, tol_rel, what, reuse_cache=False, check_free_running=True,
                          floor_factor=2.0):
    """Runs the oracle greedily (its ids are the teacher), then the engine with every step forced to the oracle's token.

    Asserted at every step of every row:
      * ``max|dlogit| <= max(tol_rel, floor_factor x floor) x logit scale`` where ``floor`` is the MEASURED distance, on the same
        trajectory and in the same norm, between two evaluations of the reference arithmetic that differ only in the fp32
        summation order of their GEMMs (fp32 SGEMM vs the cuBLAS tensor-core GEMM the reference's nn.Linear runs);
      * engine argmax == oracle argmax whenever the oracle's top-2 margin exceeds twice the measured |dlogit| of that row.
    Returns (statistics for the test log, oracle ids)."""
    dev = model.device
    T = prompts.shape[1]
    o_ids, o_scores = orc.generate(prompts, img, n_new)
    forced = o_ids[:, T:].contiguous()
    alt_scores = oracle_forced_scores(orc, prompts, img, forced, "cublas")
    out = model.generate(prompts.to(dev), img_embeds=None if img is None else img.to(dev), max_new_tokens=n_new, suppress_eos=True,
                         forced_tokens=forced, return_dict_in_generate=True, output_scores=True, reuse_cache=reuse_cache)
    own = out.sequences[:, T:]
    eos = model.cfg.eos_token_id
    B = prompts.shape[0]
    stats = dict(what=what, B=B, T=T, steps=n_new, tol_rel=tol_rel, checked=0, sub_margin=0, mismatch_sub_margin=0, max_rel_err=0.0,
                 worst_step=-1, rms_rel_err=0.0, noise_floor_rel=0.0, frac_within_tol=0.0)
    failures = []
    sq = 0.0
    for s in range(n_new):
        a = o_scores[s].float()
        b = out.scores[s].float()
        scale = a.abs().max().item()
        err_row = (a - b).abs().max(-1).values                       # [B]
        rel = (err_row.max() / scale).item()
        floor = ((a - alt_scores[s].float()).abs().max() / scale).item()
        stats["noise_floor_rel"] = max(stats["noise_floor_rel"], floor)
        sq += ((a - b) / scale).pow(2).mean().item()
        stats["frac_within_tol"] += float((err_row / scale <= tol_rel).float().mean()) / n_new
        if rel > stats["max_rel_err"]:
            stats["max_rel_err"], stats["worst_step"] = rel, s
        a_sel = a.clone()
        a_sel[:, eos] = -float("inf")                                 # both sides select with EOS suppressed
        top = a_sel.topk(2, dim=-1).values
        margin = top[:, 0] - top[:, 1]
        decisive = margin > 2 * err_row
        same = own[:, s] == forced[:, s]
        bad = decisive & ~same
        if bool(bad.any()):
            failures.append(f"step {s} rows {bad.nonzero().flatten().tolist()}: engine argmax differs from the oracle's although the "
                            f"oracle margin {margin[bad].min().item():.4g} exceeds 2 x |dlogit| {err_row[bad].max().item():.4g}")
        stats["checked"] += int(decisive.sum())
        stats["sub_margin"] += int((~decisive).sum())
        stats["mismatch_sub_margin"] += int((~decisive & ~same).sum())
    stats["rms_rel_err"] = (sq / n_new) ** 0.5
    stats["bound_rel"] = max(tol_rel, floor_factor * stats["noise_floor_rel"])
    stats["teacher_forced_token_agreement"] = float((own == forced).float().mean())
    if check_free_running and not reuse_cache:
        free = model.generate(prompts.to(dev), img_embeds=None if img is None else img.to(dev), max_new_tokens=n_new, suppress_eos=True)
        eq = (free[:, T:] == forced)
        first_div = torch.where(eq.all(-1), n_new, (~eq).float().argmax(-1))
        stats["free_running_rows_identical"] = int(eq.all(-1).sum())
        stats["free_running_first_divergence_min"] = int(first_div.min())
        stats["free_running_token_agreement"] = float(eq.float().mean())
    assert not failures, f"{what}: " + "; ".join(failures[:4]) + f" | {stats}"
    assert stats["max_rel_err"] <= stats["bound_rel"], (f"{what}: max|dlogit| = {stats['max_rel_err']:.3e} x logit scale at step "
                                                        f"{stats['worst_step']} exceeds {stats['bound_rel']:.3e} | {stats}")
    assert stats["rms_rel_err"] <= tol_rel / 2, f"{what}: rms logit error {stats['rms_rel_err']:.3e} x scale | {stats}"
    return stats, o_ids
