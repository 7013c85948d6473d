"""Shared parity helpers of the GPU tests."""
import torch


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
