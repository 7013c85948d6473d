"""GPU parity of Blip2Qformer.forward_image (native vision engine) against the CPU oracle and the reference fixture.

The reference runs this stage in fp32; the B200 path computes in fp16/bf16 with fp32 accumulation, so
the bar is the north_star tolerance: 1e-2 relative to the output scale (measured error is reported in the assert)."""
import os

import numpy as np
import pytest
import torch

from radialog_b200 import _lib, synth
from radialog_b200.vision import Blip2Qformer
from oracle import radialog_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tiny_vision_vs_oracle(cuda_dev, dtype):
    cfg = synth.tiny_vision_cfg()
    sd = synth.make_vision_weights(cfg, seed=0)
    imgs = synth.make_images(3, size=cfg.image_size, seed=1234)
    o_q, o_e = O.forward_image(imgs, sd, cfg)
    model = Blip2Qformer.from_state_dict(cfg, sd, torch_dtype=dtype, device=cuda_dev, max_batch=4)
    q, e = model.forward_image(imgs.to(cuda_dev))
    tol = 1e-2 if dtype == torch.float16 else 5e-2
    assert torch.isfinite(q).all() and torch.isfinite(e).all()
    assert rel_err(e.cpu(), o_e) <= tol, f"image_embeds rel err {rel_err(e.cpu(), o_e):.3e}"
    assert rel_err(q.cpu(), o_q) <= tol, f"q_out rel err {rel_err(q.cpu(), o_q):.3e}"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_full_resnet50_qformer_vs_reference_golden(cuda_dev, golden_dir, dtype, capsys):
    """448x448, ResNet-50 [3,4,6,3], 12-layer Q-Former: outputs of the reference's own biovil_t + Qformer modules.
    fp16 meets north_star's 1e-2 (measured 3e-3) and is the dtype bench.py / ReportPipeline run this stage in, whatever the
    LLM dtype.  bf16 (8 mantissa bits through 53 convolutions + 12 Q-Former layers) measures 2.3e-2: supported, held to
    3e-2 here, and NOT used where the 1e-2 bar applies."""
    tol = 1e-2 if dtype == torch.float16 else 3e-2
    z = np.load(os.path.join(golden_dir, "vision_r50_448.npz"))
    cfg = synth.VisionCfg(image_size=int(z["image_size"]))
    sd = synth.make_vision_weights(cfg, seed=int(z["seed"]))
    B = int(z["B"])
    imgs = synth.make_images(B, size=cfg.image_size, seed=int(z["img_seed"]))
    model = Blip2Qformer.from_state_dict(cfg, sd, torch_dtype=dtype, device=cuda_dev, max_batch=B)
    q, e = model.forward_image(imgs.to(cuda_dev))
    ref_q = torch.from_numpy(z["q_out"])
    ref_e = torch.from_numpy(z["image_embeds_sub"])
    with capsys.disabled():
        print(f"\n[vision full size {dtype}] image_embeds rel err {rel_err(e.cpu()[:, ::7, ::11], ref_e):.3e}, q_out rel err {rel_err(q.cpu(), ref_q):.3e}")
    assert rel_err(e.cpu()[:, ::7, ::11], ref_e) <= tol, f"image_embeds rel err {rel_err(e.cpu()[:, ::7, ::11], ref_e):.3e}"
    assert rel_err(q.cpu(), ref_q) <= tol, f"q_out rel err {rel_err(q.cpu(), ref_q):.3e}"
    # chunked execution (max_batch 1) against one batch of B: the GEMM tiling (token-tile width, split-K, CTA pairs) is chosen
    # per problem size, so the fp32 summation order - not the arithmetic - differs; both must meet the same bar, and agree with
    # each other to within it
    model1 = Blip2Qformer.from_state_dict(cfg, sd, torch_dtype=dtype, device=cuda_dev, max_batch=1)
    q1, _ = model1.forward_image(imgs.to(cuda_dev))
    assert rel_err(q1.cpu(), ref_q) <= tol, f"chunked q_out rel err {rel_err(q1.cpu(), ref_q):.3e}"
    assert rel_err(q1.cpu(), q.cpu()) <= tol


def test_two_image_temporal_branch_tiny_vs_oracle(cuda_dev):
    """biovil_t/encoder.py:117-123 + VisionTransformerPooler (transformer.py:28-118) through rd_vision_forward_temporal.  The
    previous image moves the output only a little, so besides the usual tolerance the CHANGE it causes (two-image minus
    single-image output) must itself match the oracle's change."""
    cfg = synth.tiny_vision_cfg()
    sd = synth.make_vision_weights(cfg, seed=0)
    cur = synth.make_images(3, size=cfg.image_size, seed=1234)
    prev = synth.make_images(3, size=cfg.image_size, seed=4242)
    o_q2, o_e2 = O.forward_image(cur, sd, cfg, prev)
    o_q1, o_e1 = O.forward_image(cur, sd, cfg)
    model = Blip2Qformer.from_state_dict(cfg, sd, torch_dtype=torch.float16, device=cuda_dev, max_batch=2)     # 3 images: chunked 2 + 1
    q2, e2 = model.forward_image(cur.to(cuda_dev), prev.to(cuda_dev))
    q1, e1 = model.forward_image(cur.to(cuda_dev))
    assert rel_err(e2.cpu(), o_e2) <= 1e-2 and rel_err(q2.cpu(), o_q2) <= 1e-2
    assert rel_err(e1.cpu(), o_e1) <= 1e-2 and rel_err(q1.cpu(), o_q1) <= 1e-2
    d_ref, d_own = (o_e2 - o_e1), (e2 - e1).cpu()
    assert d_ref.abs().max() > 1e-3
    assert rel_err(d_own, d_ref) <= 0.25, f"effect of the previous image differs: {rel_err(d_own, d_ref):.3e}"
    with pytest.raises(AssertionError):
        model.forward_image(cur.to(cuda_dev), prev[:2].to(cuda_dev))


def test_two_image_temporal_branch_full_size_vs_reference_golden(cuda_dev, golden_dir):
    """Full ResNet-50 + 3-block pooler (dim 256, 8 heads, 2 x 196 tokens) against the output of the reference's own
    MultiImageEncoder / VisionTransformerPooler modules (tests/golden/vision_temporal_r50_448.npz)."""
    z = np.load(os.path.join(golden_dir, "vision_temporal_r50_448.npz"))
    cfg = synth.VisionCfg(image_size=int(z["image_size"]))
    sd = synth.make_vision_weights(cfg, seed=int(z["seed"]))
    B = int(z["B"])
    cur = synth.make_images(B, size=cfg.image_size, seed=int(z["img_seed"]))
    prev = synth.make_images(B, size=cfg.image_size, seed=int(z["prev_seed"]))
    model = Blip2Qformer.from_state_dict(cfg, sd, torch_dtype=torch.float16, device=cuda_dev, max_batch=B)
    q, e = model.forward_image(cur.to(cuda_dev), prev.to(cuda_dev))
    ref_q, ref_e = torch.from_numpy(z["q_out"]), torch.from_numpy(z["image_embeds_sub"])
    assert rel_err(e.cpu()[:, ::7, ::11], ref_e) <= 1e-2, f"image_embeds rel err {rel_err(e.cpu()[:, ::7, ::11], ref_e):.3e}"
    assert rel_err(q.cpu(), ref_q) <= 1e-2, f"q_out rel err {rel_err(q.cpu(), ref_q):.3e}"
    q1, _ = model.forward_image(cur.to(cuda_dev))
    assert (q - q1).abs().max().item() > 1e-3


def test_image_shape_is_validated(cuda_dev):
    cfg = synth.tiny_vision_cfg()
    sd = synth.make_vision_weights(cfg, seed=0)
    model = Blip2Qformer.from_state_dict(cfg, sd, device=cuda_dev, max_batch=2)
    with pytest.raises(ValueError):
        model.forward_image(torch.zeros(1, 1, cfg.image_size, cfg.image_size))


def test_image_to_report_end_to_end(cuda_dev):
    """forward_image -> splice -> greedy decode through ReportPipeline equals the oracle chain on the same inputs
    (tie rule as in test_gpu_llm)."""
    from radialog_b200.llm import LlamaForCausalLM
    from radialog_b200.pipeline import ReportPipeline
    from parity_util import assert_ids_match
    vcfg = synth.tiny_vision_cfg(q_hidden=768, q_heads=12, q_intermediate=256, joint_feature_size=128)
    lcfg = synth.tiny_llama_cfg()
    vsd = synth.make_vision_weights(vcfg, seed=0)
    lsd = {k: v.to(torch.float16).float() for k, v in synth.make_llama_weights(lcfg, seed=0, dtype=torch.float32).items()}
    vis = Blip2Qformer.from_state_dict(vcfg, vsd, device=cuda_dev, max_batch=4)
    llm = LlamaForCausalLM.from_state_dict(lcfg, lsd, device=cuda_dev)
    pipe = ReportPipeline(vis, llm)
    imgs = synth.make_images(4, size=vcfg.image_size, seed=1234)
    prompts = synth.make_prompts(4, seed=4321, ragged=True)
    out = pipe.generate(imgs.to(cuda_dev), prompts.to(cuda_dev), max_new_tokens=8, suppress_eos=True).cpu()
    orc = O.LlamaOracle(lcfg, lsd, torch.float16)
    o_q, _ = O.forward_image(imgs, vsd, vcfg)
    o_ids, o_scores = orc.generate(prompts, o_q, 8, suppress_eos=True, return_scores=True)
    assert_ids_match(out, o_ids, o_scores, prompts.shape[1], torch.float16, "image->report", min_exact_rows=0.5)
