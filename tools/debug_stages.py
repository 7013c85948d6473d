#!/usr/bin/env python
"""Staged bring-up on the GPU box: every stage runs in its own subprocess under a timeout, so a hang or a trap in one
stage is reported and the others still run.  python tools/debug_stages.py [stage ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = {}


def stage(fn):
    STAGES[fn.__name__] = fn
    return fn


def _lin(M, N, K, algo, dtype_name="float16", act=0, splits=0, check=True):
    import ctypes as C
    import torch
    from radialog_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    dtype = getattr(torch, dtype_name)
    g = torch.Generator().manual_seed(M + N + K)
    rows = 2 * N if act == 3 else N
    x = (torch.randn(M, K, generator=g) * 0.5).to(dtype).to(dev)
    w = (torch.randn(rows, K, generator=g) * 0.05).to(dtype).to(dev)
    out = torch.zeros(M, N, device=dev, dtype=dtype)
    ws = torch.zeros(64 << 20, dtype=torch.uint8, device=dev)
    e = _lib.Epilogue()
    e.act = act
    e.res_mode = 1
    lib.rd_linear_force_splits(splits)
    st = lib.rd_linear(x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, C.byref(e), _lib.dtype_code(dtype), algo,
                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rd_linear")
    torch.cuda.synchronize()
    if act == 3:
        gt = (x.float() @ w[:N].float().t()).to(dtype)
        ut = (x.float() @ w[N:].float().t()).to(dtype)
        ref = torch.nn.functional.silu(gt.float()).to(dtype) * ut
    else:
        ref = (x.float() @ w.float().t()).to(dtype)
    err = (out.float() - ref.float()).abs().max().item()
    print(f"  M={M} N={N} K={K} algo={algo} act={act} splits={splits} {dtype_name}: max abs err {err:.4g} (ref max {ref.float().abs().max().item():.3g})", flush=True)
    if check:
        assert err <= 0.02 * ref.float().abs().max().item() + 1e-3, "MISMATCH"


@stage
def s01_device():
    import torch
    from radialog_b200 import _lib
    lib = _lib.load()
    print("  device:", torch.cuda.get_device_name(0), "ok:", lib.rd_device_ok(0), flush=True)


@stage
def s02_gemv():
    _lin(1, 256, 256, 1)
    _lin(4, 1001, 704, 1)
    _lin(2, 512, 512, 1, act=3)


@stage
def s03_simt():
    _lin(48, 384, 512, 3)
    _lin(48, 384, 512, 3, act=3)


@stage
def s04_tc_one_tile():
    _lin(16, 128, 64, 2)


@stage
def s05_tc_k256():
    _lin(16, 128, 256, 2)
    _lin(32, 256, 512, 2)


@stage
def s06_tc_shapes():
    _lin(64, 384, 704, 2)
    _lin(100, 1001, 704, 2)
    _lin(300, 256, 152, 2)
    _lin(1000, 768, 3072, 2)
    _lin(32, 256, 512, 2, dtype_name="bfloat16")


@stage
def s07_tc_splitk():
    _lin(32, 512, 4096, 2, splits=3)
    _lin(32, 512, 4096, 2, splits=7)


@stage
def s08_tc_swiglu():
    _lin(32, 256, 512, 2, act=3)
    _lin(300, 704, 256, 2, act=3)
    _lin(32, 512, 4096, 2, act=3, splits=3)


def _tiny_llm(algo, graph, B=3, new=6):
    import torch
    from radialog_b200 import synth
    from radialog_b200.llm import LlamaForCausalLM
    from oracle import radialog_oracle as O
    dev = torch.device("cuda:0")
    cfg = synth.tiny_llama_cfg()
    sd = {k: v.to(torch.float16).float() for k, v in synth.make_llama_weights(cfg, seed=0, dtype=torch.float32).items()}
    m = LlamaForCausalLM.from_state_dict(cfg, sd, device=dev)
    m.set_algo(algo)
    m.use_cuda_graph = graph
    prompts = synth.make_prompts(B, seed=4321, ragged=True)
    img = torch.randn(B, 32, cfg.qformer_hidden, generator=torch.Generator().manual_seed(99)) * 0.5
    orc = O.LlamaOracle(cfg, sd, torch.float16)
    mask = prompts.ne(0).long()
    ol, _ = orc.forward(prompts, mask, orc.positions_from_mask(mask), None, img)
    print("  prefill_logits ...", flush=True)
    lg = m.prefill_logits(prompts.to(dev), img.to(dev)).cpu()
    err = (lg.float() - ol.float()).abs()[mask.bool()].max().item()
    print(f"  prefill logits max abs err {err:.4g} (scale {ol.float().abs().max().item():.3g})", flush=True)
    print("  generate ...", flush=True)
    ids = m.generate(prompts.to(dev), img_embeds=img.to(dev), max_new_tokens=new).cpu()
    oi = orc.generate(prompts, img, new)
    print("  ids equal:", bool(ids.shape == oi.shape and (ids == oi).all()), ids[:, -new:].tolist(), oi[:, -new:].tolist(), flush=True)


@stage
def s09_llm_simt_eager():
    _tiny_llm(3, False)


@stage
def s10_llm_auto_eager():
    _tiny_llm(0, False)


@stage
def s11_llm_auto_graph():
    _tiny_llm(0, True, new=10)


@stage
def s12_llm_b32_graph():
    _tiny_llm(0, True, B=32, new=8)


@stage
def s13_vision_tiny():
    import torch
    from radialog_b200 import synth
    from radialog_b200.vision import Blip2Qformer
    from oracle import radialog_oracle as O
    dev = torch.device("cuda:0")
    cfg = synth.tiny_vision_cfg()
    sd = synth.make_vision_weights(cfg, seed=0)
    imgs = synth.make_images(3, size=cfg.image_size, seed=1234)
    oq, oe = O.forward_image(imgs, sd, cfg)
    m = Blip2Qformer.from_state_dict(cfg, sd, device=dev, max_batch=4)
    q, e = m.forward_image(imgs.to(dev))
    torch.cuda.synchronize()
    print(f"  embeds rel err {((e.cpu() - oe).abs().max() / oe.abs().max()).item():.3e}  q rel err {((q.cpu() - oq).abs().max() / oq.abs().max()).item():.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or sorted(STAGES)
    for n in names:
        t = time.time()
        print(f"== {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], timeout=150, capture_output=True, text=True)
            print(r.stdout[-3000:], end="")
            if r.returncode != 0:
                print(f"   FAILED rc={r.returncode}\n{r.stderr[-2500:]}")
            else:
                print(f"   ok ({time.time() - t:.1f}s)")
        except subprocess.TimeoutExpired as ex:
            print(f"   TIMEOUT after 150 s\n{(ex.stdout or b'')[-2000:]}\n{(ex.stderr or b'')[-2000:]}")
        sys.stdout.flush()
