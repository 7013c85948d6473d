#!/usr/bin/env python
"""bench.py — image->report throughput of the B200-native RaDialog hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                     # own arm (torchrun-launched for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W    # the reference algorithm's CPU path (oracle port)
    python bench.py --workload chat                                   # BASELINE.json configs[4]: multi-turn, p50 per-turn latency

A "step" is one pass of the whole hot path over one batch of synthetic input per GPU: B images (448x448) ->
BioViL-T ResNet-50 -> Q-Former -> splice into a T=64 prompt -> Vicuna-7B prefill -> 128 greedy tokens.  Workload =
BASELINE.json configs[2] (batch 32 per GPU; configs[3] is the same shard at N=8: 256 images).  Weights are random-init
(seeded) at the real architecture sizes; EOS is suppressed so every step does the same work.  The LLM computes in bf16
(configs[2]); the vision stage in fp16 (the dtype that meets the 1e-2 parity bar at full size, tests/test_gpu_vision.py).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import csv
import gzip
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "reports_per_sec"
UNIT = "reports/s"
T_PROMPT = 64
W_DEC_BYTES = 13_214_695_424          # decode-step weight bytes (V=32001, 2-byte weights)
KV_BYTES_PER_TOKEN = 524_288           # per sequence per cached token, 32 layers
GATE_UP_BYTES = 2 * 2 * 11008 * 4096   # fused gate|up weights per launch (per-op path)
LAYER_W_BYTES = 32 * 202_383_360 * 2   # the 32 decoder layers' weights, 2-byte (12,952,535,040 B)
LAYER_MM_PARAMS = 202_375_168          # matmul parameters per decoder layer (SURVEY.md 8d prefill FLOPs)
VISION_GFLOP_PER_IMAGE = 33.96 + 11.1  # image encoder + Q-Former
PROFILE_STEPS = 8
NCU_GATE_UP_CSV = os.path.join(ROOT, "profiles", "ncu_r2_gate_up_raw.csv.gz")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d.get("hbm_gbs", 6650.0), "tf_burst": d.get("bf16_tflops", 1590.0), "tf_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the gate|up decode GEMM, parsed from the committed raw page
    of the ncu --set full capture (profiles/); None if the capture is not there."""
    if not os.path.exists(NCU_GATE_UP_CSV):
        return None, "no ncu capture committed"
    try:
        with gzip.open(NCU_GATE_UP_CSV, "rt") as f:
            rows = list(csv.reader(f))
        hdr = rows[0]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        units = rows[1]
        mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for r in rows[2:]:
            if len(r) > max(ir, iw) and "linear_tc_kernel" in r[ik]:
                vals.append(float(r[ir].replace(",", "")) * mul.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * mul.get(units[iw], 1.0))
        if not vals:
            return None, "no linear_tc_kernel rows in the ncu capture"
        return sum(vals) / len(vals), f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(vals)} launches ({os.path.relpath(NCU_GATE_UP_CSV, ROOT)})"
    except Exception as e:          # a malformed capture must not take the bench down
        return None, f"ncu capture unreadable: {e}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference algorithm on the host cores, REAL 32-layer model
# ----------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference algorithm (oracle/radialog_oracle.py: fp32, what the reference computes on CPU) at the real size:
    full ResNet-50 + Q-Former and all 32 Vicuna-7B layers (27 GB of fp32 weights on the host), all host threads.
    No layer extrapolation: every timing runs the whole model.  `sample(n_dec)` = 1 image -> Q-Former -> T=64 prefill ->
    n_dec greedy tokens (reference loop: all-position lm_head, torch.cat KV cache)."""

    def __init__(self):
        from oracle import radialog_oracle as O
        from radialog_b200 import synth
        self.O = O
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        t0 = time.time()
        self.vcfg = synth.VisionCfg()
        self.vsd = synth.make_vision_weights(self.vcfg, seed=0)
        self.img = synth.make_images(1, seed=1234)
        self.prompts = synth.make_prompts(1, seed=4321)
        cfg = synth.LlamaCfg()
        if torch.cuda.is_available():          # same seeded generator as the own arm; RNG on the GPU, tensors moved to the host
            sd = synth.make_llama_weights(cfg, seed=0, dtype=torch.float32, lora=True, device="cuda")
            sd = {k: v.cpu() for k, v in sd.items()}
            torch.cuda.empty_cache()
        else:
            sd = synth.make_llama_weights(cfg, seed=0, dtype=torch.float32, lora=True)
        self.orc = O.LlamaOracle(cfg, sd, torch.float32)
        self.setup_s = time.time() - t0

    @torch.no_grad()
    def sample(self, n_dec: int):
        O, orc = self.O, self.orc
        t0 = time.perf_counter()
        q, _ = O.forward_image(self.img, self.vsd, self.vcfg)
        t1 = time.perf_counter()
        ids = self.prompts
        mask = ids.ne(0).long()
        logits, past = orc.forward(ids, mask, orc.positions_from_mask(mask), None, q)
        ids = torch.cat([ids, logits[:, -1].argmax(-1)[:, None]], -1)
        t2 = time.perf_counter()
        for _ in range(n_dec):
            mask = torch.cat([mask, mask.new_ones(1, 1)], -1)
            pos = orc.positions_from_mask(mask)
            logits, past = orc.forward(ids[:, -1:], mask, pos[:, -1:], past, None)
            ids = torch.cat([ids, logits[:, -1].argmax(-1)[:, None]], -1)
        t3 = time.perf_counter()
        return {"vision_s": t1 - t0, "prefill_s": t2 - t1, "decode_s_per_token": (t3 - t2) / max(1, n_dec), "wall_s": t3 - t0, "n_dec": n_dec}

    @staticmethod
    def report_seconds(s, new_tokens):
        """Seconds per report of `new_tokens` tokens from one sample: the first token comes out of the prefill, the other
        new_tokens-1 each cost one (measured, full-depth) decode step."""
        return s["vision_s"] + s["prefill_s"] + (new_tokens - 1) * s["decode_s_per_token"]


def cpu_line(samples, ref, new_tokens, full_report_s=None):
    s_rep = sum(CpuReference.report_seconds(s, new_tokens) for s in samples) / len(samples)
    mean = lambda k: sum(s[k] for s in samples) / len(samples)
    sample = (f"{len(samples)} x [1 image through the full ResNet-50 + Q-Former (fp32) + the full 32-layer Vicuna-7B (fp32, 27 GB of weights): "
              f"T={T_PROMPT} prefill + {samples[0]['n_dec']} greedy decode steps]; seconds per {new_tokens}-token report = vision {mean('vision_s'):.3f} s + "
              f"prefill {mean('prefill_s'):.2f} s + {new_tokens - 1} x decode {mean('decode_s_per_token'):.3f} s/token (every figure measured at full depth; "
              f"only the token count is scaled)")
    out = {"value": 1.0 / s_rep, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": sample, "s_per_report": s_rep}
    if full_report_s is not None:
        out["full_report_s_measured"] = full_report_s
        out["value_from_full_report"] = 1.0 / full_report_s
    return out


def run_reference_arm(args):
    """Each of the K timed steps is one bounded sample (full-depth model, a few decode tokens).  When --warmup >= 1 the first
    warm-up step is one REAL complete report (all `new_tokens` tokens), reported beside the scaled figure."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    full = None
    n_dec = args.ref_dec_tokens
    t_start = time.time()
    for i in range(args.warmup):
        if i == 0 and not args.no_full_report:
            s = ref.sample(args.new_tokens - 1)
            full = s["wall_s"]
        else:
            ref.sample(n_dec)
    samples = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        samples.append(ref.sample(n_dec))
    wall = time.perf_counter() - t0
    cb = cpu_line(samples, ref, args.new_tokens, full)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (seeded random-init weights at Vicuna-7B / ResNet-50 / Q-Former sizes)",
            "config": {"workload": f"configs[2]: batch={args.batch} images/GPU, 448x448 -> BioViL-T ResNet-50 -> Q-Former -> Vicuna-7B, T={T_PROMPT} prompt, {args.new_tokens} greedy tokens (EOS suppressed)",
                       "note": "reference arm = the reference algorithm's CPU path (oracle port, torch fp32, all host threads) on the REAL 32-layer model; "
                               "a step = one bounded sample of the workload (one report's vision + prefill + a few decode steps); "
                               "ms_per_step is the wall time of a step; value = reports/s of a full 128-token report derived from the measured phases",
                       "setup_s": round(ref.setup_s, 1), "total_s": round(time.time() - t_start + ref.setup_s, 1)},
            "cpu_baseline": {k: v for k, v in cb.items() if k != "s_per_report"},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# reference GPU PyTorch path (denominator of north_star's ">= 10x at batch 32"): the oracle's ops on cuda, cuBLAS GEMMs
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def reference_gpu_path(lsd, lcfg, dtype, dev, prompts, img_tokens, new_tokens):
    """The reference's GPU execution restated: eager PyTorch ops, three separate q/k/v Linears + unmerged
    LoRA, torch.cat KV cache, all-position lm_head, per-token Python loop, cuBLAS GEMMs in the model dtype.  transformers
    4.28.1 `generate` itself is not installable offline; the loop is oracle/radialog_oracle.py's faithful restatement."""
    import torch.nn.functional as F
    from oracle import radialog_oracle as O
    old = O._mm
    O._mm = lambda x, w, dt: F.linear(x, w)
    try:
        orc = O.LlamaOracle(lcfg, lsd, dtype)
        orc.cos, orc.sin = orc.cos.to(dev), orc.sin.to(dev)

        def run(n):
            with torch.device(dev):
                return orc.generate(prompts, img_tokens, n, suppress_eos=True)

        run(4)                                  # warm-up: cuBLAS heuristics, allocator
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(new_tokens)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    finally:
        O._mm = old


# ----------------------------------------------------------------------------------------------------------------
# own arm
# ----------------------------------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch.distributed as dist
    from radialog_b200 import _lib, synth
    from radialog_b200.llm import LlamaForCausalLM
    from radialog_b200.vision import Blip2Qformer
    from radialog_b200.pipeline import ReportPipeline, broadcast_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a B200; there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    assert lib.rd_device_ok(local), lib.rd_last_error().decode()
    dtype = {"float16": torch.float16, "bfloat16": torch.bfloat16}[args.dtype]
    vis_dtype = torch.float16
    B, NEW = args.batch, args.new_tokens

    # ---- weights: built once on rank 0 (GPU RNG for the 7B model), one NCCL broadcast at load --------------------------------
    lcfg, vcfg = synth.LlamaCfg(), synth.VisionCfg()
    t0 = time.time()
    lsd = synth.make_llama_weights(lcfg, seed=0, dtype=dtype, device=str(dev)) if rank == 0 else None
    vsd = synth.make_vision_weights(vcfg, seed=0) if rank == 0 else None
    if world > 1:
        lsd = broadcast_state_dict(lsd, src=0, device=dev)
        vsd = broadcast_state_dict({k: v.to(dev) for k, v in vsd.items()} if rank == 0 else None, src=0, device=dev)
        vsd = {k: v.cpu() for k, v in vsd.items()}
    llm = LlamaForCausalLM.from_state_dict(lcfg, lsd, torch_dtype=dtype, device=dev)
    want_ref_gpu = rank == 0 and world == 1 and not args.no_ref_gpu
    if not want_ref_gpu:
        del lsd
    torch.cuda.empty_cache()
    vis = Blip2Qformer.from_state_dict(vcfg, vsd, torch_dtype=vis_dtype, device=dev, max_batch=B)
    pipe = ReportPipeline(vis, llm)
    llm.reserve(B, T_PROMPT + NEW + 2)
    lib.rd_set_pdl(1 if args.pdl else 0)
    t_load = time.time() - t0

    # ---- inputs: per-rank shard of the global batch (weak scaling: B per GPU) --------------------------------------------------
    imgs_host = synth.make_images(B, seed=1234 + rank).pin_memory()
    prompts_host = synth.make_prompts(B, seed=4321 + rank).pin_memory()
    imgs_dev, prompts_dev = imgs_host.to(dev), prompts_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return pipe.generate(imgs_dev, prompts_dev, max_new_tokens=NEW, suppress_eos=True)

    def step_e2e():
        i = imgs_host.to(dev, non_blocking=True)
        p = prompts_host.to(dev, non_blocking=True)
        return pipe.generate(i, p, max_new_tokens=NEW, suppress_eos=True).cpu()

    def launches():
        return llm.launch_count() + vis.launch_count()

    for _ in range(args.warmup):
        step_resident()
    stats = []
    barrier()
    l0 = launches()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_resident()
            stats.append(dict(pipe.last_stats))
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    n_launch = launches() - l0
    # end-to-end through the public API with host buffers
    step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        seqs = step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        nl = torch.tensor([float(n_launch)], device=dev, dtype=torch.float64)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)          # whole-job launch count
        n_launch = int(nl.item())

    # ---- roofline of the dominant kernel + whole decode step + prefill / vision fractions (rank 0) ------------------------------
    out = None
    if rank == 0:
        pk = peaks()
        hbm_peak = pk["hbm"]
        roof, prof = dominant_kernel_roofline(llm, B, hbm_peak, pk["src"])
        dec_ms = sum(s["decode_ms"] for s in stats) / len(stats) / (NEW - 1)
        c_mid = T_PROMPT + NEW // 2
        step_bytes = W_DEC_BYTES + B * KV_BYTES_PER_TOKEN * (c_mid + 1)
        step_gbs = step_bytes / (dec_ms * 1e-3) / 1e9
        value = B * world * args.steps / (ms * 1e-3)
        phases = {k: sum(s[k] for s in stats) / len(stats) for k in ("vision_ms", "prefill_ms", "decode_ms")}
        # the dominant kernel's share of a decode step (32 launches per step), to set beside the ncu launch list's share
        roof["share_of_decode_step"] = roof["ms_per_launch"] * lcfg.num_hidden_layers / dec_ms
        pre_flop = (2.0 * B * T_PROMPT * lcfg.num_hidden_layers * LAYER_MM_PARAMS + 2.0 * B * lcfg.hidden_size * lcfg.vocab_size
                    + 4.0 * B * lcfg.num_attention_heads * 128 * T_PROMPT * (T_PROMPT + 1) / 2 * lcfg.num_hidden_layers)
        pre_tf = pre_flop / (phases["prefill_ms"] * 1e-3) / 1e12
        vis_tf = VISION_GFLOP_PER_IMAGE * 1e9 * B / (phases["vision_ms"] * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if dtype == torch.float16 else "bf16", "data": "synthetic (seeded random-init weights at Vicuna-7B / ResNet-50 / Q-Former sizes)",
            "config": {"workload": f"configs[2]: batch={B} images/GPU, 448x448 -> BioViL-T ResNet-50 -> Q-Former -> Vicuna-7B, T={T_PROMPT} prompt, {NEW} greedy tokens (EOS suppressed)",
                       "parallelism": f"dp{world} (one weight broadcast at load, no data-path collective)", "global_batch": B * world,
                       "l2": "decode streams 13.2 GB of weights per step (>> 126 MB L2), so no L2 flush is needed between steps",
                       "vision_dtype": "f16", "seeds": {"weights": 0, "images": 1234, "prompts": 4321}, "pdl": bool(args.pdl), "load_s": round(t_load, 1)},
            "clocks": clk.summary(),
            "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": imgs_host.numel() * 4 + prompts_host.numel() * 8,
                    "d2h_bytes_per_step": int(seqs.numel() * 8)},
            "gpu_launches": int(n_launch),
            "roofline": roof,
            "decode_step": {"ms": dec_ms, "algorithmic_bytes": step_bytes, "achieved_gbs": step_gbs, "frac_of_hbm_peak": step_gbs / hbm_peak,
                            "tokens_per_s": B / (dec_ms * 1e-3)},
            "prefill": {"ms": phases["prefill_ms"], "algorithmic_tflop": pre_flop / 1e12, "achieved_tflops": pre_tf, "peak_tflops": pk["tf_sustained"],
                        "frac_of_bf16_sustained": pre_tf / pk["tf_sustained"], "bound": "tensor"},
            "vision": {"ms": phases["vision_ms"], "algorithmic_gflop_per_image": VISION_GFLOP_PER_IMAGE, "achieved_tflops": vis_tf,
                       "frac_of_bf16_sustained": vis_tf / pk["tf_sustained"], "bound": "tensor (launch-bound in practice)"},
            "phases_ms": phases,
            "kernel_classes_ms_per_decode_step": {k: v["ms"] / PROFILE_STEPS for k, v in prof.items()},
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_b1 and world == 1:
            out["b1"] = bench_b1(pipe, llm, dev, NEW, peaks()["hbm"])
        if want_ref_gpu:
            g = torch.Generator().manual_seed(7)
            q_out, _ = vis.forward_image(imgs_dev)
            ms_ref = reference_gpu_path(lsd, lcfg, dtype, dev, prompts_dev, q_out, NEW)
            own_llm_ms = out["phases_ms"]["prefill_ms"] + out["phases_ms"]["decode_ms"]
            out["reference_gpu"] = {"what": "reference GPU PyTorch path restated (oracle ops on cuda: eager, cuBLAS GEMMs in the model dtype, torch.cat KV cache, "
                                            "all-position lm_head, per-token Python loop), LLM part (prefill + 128 greedy tokens) of the same batch on the same GPU",
                                    "ms_per_batch": ms_ref, "reports_per_s": B * 1e3 / ms_ref, "ms_per_decode_step": ms_ref / NEW,
                                    "own_ms_per_batch_llm_part": own_llm_ms, "speedup": ms_ref / own_llm_ms, "target": 10.0}
            del lsd
            torch.cuda.empty_cache()
        if not args.no_cpu and world == 1:
            ref = CpuReference()
            ref.sample(1)                                   # warm-up (thread pool, allocator)
            cb = cpu_line([ref.sample(args.ref_dec_tokens * 2)], ref, NEW)
            cb.pop("s_per_report")
            out["cpu_baseline"] = cb
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(llm, B, hbm_peak, peak_src):
    """Dominant kernel of the default (one kernel per op) decode path = the fused gate|up projection GEMM
    (linear_tc_kernel<NT, SWIGLU>, ~28 % of a decode step's kernel time, profiles/).
    Algorithmic bytes per launch = its weights, 180,355,072 B.  Average launch duration, live, with CUDA
    events on the launching stream: the 32 layers' gate|up GEMMs (32 distinct 180 MB weight buffers = 5.8 GB >> L2, so
    every launch streams from HBM) are launched back to back through the same C-ABI entry point and launch attributes
    (programmatic dependent launch as in the timed run) the engine uses, 8 rounds, events around the whole series: an
    IN-PIPELINE figure (the next launch's prologue overlaps this one's tail, as it does inside a decode step).
    `isolated_ms_per_launch` is the same kernel bracketed by its own pair of events inside eager decode steps (includes the
    launch gap on both sides) - the pessimistic, solo-launch figure."""
    import ctypes as C
    from radialog_b200 import _lib
    lib = _lib.load()
    prof = llm.profile_decode_steps(B, T_PROMPT, steps=PROFILE_STEPS)
    iso_ms = prof["gate_up"]["ms"] / max(1, prof["gate_up"]["launches"])
    cfg, dev, dt = llm.cfg, llm.device, llm.dtype
    H, I = cfg.hidden_size, cfg.intermediate_size
    x = (torch.randn(B, H, device=dev) * 0.5).to(dt)
    out = torch.empty(B, I, device=dev, dtype=dt)
    ws = torch.zeros(int(lib.rd_linear_workspace_bytes(B, I, H)) + 256, dtype=torch.uint8, device=dev)
    e = _lib.Epilogue()
    e.act = _lib.ACT_SWIGLU
    e.res_mode = 1
    st = torch.cuda.current_stream().cuda_stream
    layers = llm.model.layers_w

    def series():
        for lw in layers:
            _lib.check(lib.rd_linear(x.data_ptr(), H, lw["gate_up"].data_ptr(), H, out.data_ptr(), I, B, I, H, C.byref(e),
                                     _lib.dtype_code(dt), _lib.ALGO_AUTO, ws.data_ptr(), ws.numel(), st), "rd_linear")

    series()
    torch.cuda.synchronize()
    rounds = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        series()
    e1.record()
    torch.cuda.synchronize()
    gu_ms = e0.elapsed_time(e1) / (rounds * len(layers))
    gu_gbs = GATE_UP_BYTES / (gu_ms * 1e-3) / 1e9 if gu_ms > 0 else 0.0
    kern = "linear_tc_kernel<NT=32,SWIGLU>" if B > 16 else "linear_tc_kernel<NT=16,SWIGLU>"
    traffic, traffic_src = ncu_traffic_per_launch() if B > 16 else (None, "not captured at this batch size")
    roof = {"kernel": f"{kern} (gate|up projection of one decoder layer, decode, B={B})", "bound": "hbm",
            "achieved": gu_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gu_gbs / hbm_peak, "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": GATE_UP_BYTES, "ms_per_launch": gu_ms,
            "timing": "in-pipeline (back-to-back launches with programmatic dependent launch, as inside a decode step)",
            "isolated_ms_per_launch": iso_ms, "isolated_frac": GATE_UP_BYTES / (iso_ms * 1e-3) / 1e9 / hbm_peak if iso_ms > 0 else None,
            "traffic_source": traffic_src,
            "how": f"CUDA events around {rounds} x {len(layers)} back-to-back launches over the 32 layers' distinct weights, launch "
                   "attributes as in the timed run; after the timed region"}
    return roof, prof


def bench_b1(pipe, llm, dev, new_tokens, hbm_peak):
    """configs[1]: batch=1 image->report latency (single-token GEMV decode)."""
    from radialog_b200 import synth
    img = synth.make_images(1, seed=1234).to(dev)
    prm = synth.make_prompts(1, seed=4321).to(dev)
    for _ in range(2):
        pipe.generate(img, prm, max_new_tokens=new_tokens, suppress_eos=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    pipe.generate(img, prm, max_new_tokens=new_tokens, suppress_eos=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    dec_ms = pipe.last_stats["decode_ms"] / (new_tokens - 1)
    step_bytes = W_DEC_BYTES + KV_BYTES_PER_TOKEN * (T_PROMPT + new_tokens // 2 + 1)
    roof, _ = dominant_kernel_roofline(llm, 1, hbm_peak, "")
    return {"reports_per_s": 1e3 / ms, "ms_per_report": ms, "decode_ms_per_token": dec_ms,
            "decode_step_gbs": step_bytes / (dec_ms * 1e-3) / 1e9, "decode_step_frac_of_hbm_peak": step_bytes / (dec_ms * 1e-3) / 1e9 / hbm_peak,
            "gate_up_gbs": roof["achieved"], "gate_up_frac": roof["frac"]}


# ----------------------------------------------------------------------------------------------------------------
# configs[4]: interactive multi-turn (demo.py:245-305): 8 conversations x (report + 4 follow-ups), p50 per-turn latency
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def run_chat(args):
    from radialog_b200 import _lib, synth
    from radialog_b200.llm import LlamaForCausalLM
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    lib = _lib.load()
    lib.rd_set_pdl(1 if args.pdl else 0)
    dtype = torch.float16                                   # the reference's chat dtype (demo.py:225)
    cfg = synth.LlamaCfg()
    sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device=str(dev))
    llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
    del sd
    B, TURNS, NEW, FOLLOW = 8, 4, 64, 24
    prompts = synth.make_prompts(B, seed=4321).to(dev)
    img = (torch.randn(B, 32, 768, generator=torch.Generator().manual_seed(7)) * 0.5).to(dev)
    g = torch.Generator().manual_seed(99)
    follows = [torch.randint(3, 32000, (B, FOLLOW), generator=g).to(dev) for _ in range(TURNS)]
    llm.reserve(B, prompts.shape[1] + (TURNS + 1) * (NEW + FOLLOW) + 8)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), r

    def conversation(reuse, teacher=None):
        """teacher: per-turn sequences of a previous run; the same conversation is replayed with every decode step forced to
        its tokens, so the two modes are compared on identical inputs at every step (a free-running near-tie would fork them)."""
        lat, seqs, agree, n_cmp = [], [], 0, 0
        conv = prompts
        for t in range(TURNS + 1):
            forced = None if teacher is None else teacher[t][:, conv.shape[1]:]
            ms, res = timed(lambda: llm.generate(conv, img_embeds=img, max_new_tokens=NEW, suppress_eos=True, reuse_cache=reuse and t > 0,
                                                 forced_tokens=forced, return_dict_in_generate=True))
            lat.append(ms)
            own = res.sequences
            if teacher is not None:
                agree += int((own[:, conv.shape[1]:] == forced).sum())
                n_cmp += forced.numel()
                own = teacher[t]
            seqs.append(own)
            if t < TURNS:
                conv = torch.cat([own, follows[t]], -1)
        return lat, seqs, (agree / n_cmp if n_cmp else None)

    conversation(True)                                      # warm-up (graph capture, lazy attribute setup)
    lat_full, seq_full, _ = conversation(False)             # reference behaviour: re-prefill the growing conversation every turn
    lat_reuse, _, agree = conversation(True, teacher=seq_full)
    lat_reuse_free, _, _ = conversation(True)
    p50 = lambda v: sorted(v)[len(v) // 2]
    line = {"metric": "p50_per_turn_latency_ms", "value": p50(lat_reuse_free[1:]), "unit": "ms", "n_gpus": 1, "steps": TURNS, "warmup": 1,
            "ms_per_step": p50(lat_reuse_free[1:]), "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic (seeded random-init weights at Vicuna-7B size)",
            "config": {"workload": f"configs[4]: {B} conversations x (report + {TURNS} follow-ups of {FOLLOW} new text ids), {NEW} new tokens per turn, KV prefix reuse (demo.py:245-305)"},
            "per_turn_ms_prefix_reuse": [round(x, 1) for x in lat_reuse_free], "per_turn_ms_full_reprefill": [round(x, 1) for x in lat_full],
            "p50_follow_up_ms_prefix_reuse": round(p50(lat_reuse_free[1:]), 1), "p50_follow_up_ms_full_reprefill": round(p50(lat_full[1:]), 1),
            "token_identity": {"how": "prefix-reuse path replayed on the full-re-prefill path's conversation with every step teacher-forced to its tokens; "
                                      "fraction of (row, step) pairs where the reuse path's own argmax equals it (full-size oracle check: "
                                      "tests/test_gpu_parity_full.py::test_config5_*)", "argmax_agreement": agree,
                               "per_turn_ms_teacher_forced": [round(x, 1) for x in lat_reuse]}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="report", choices=["report", "chat"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--dtype", default="bfloat16", choices=["float16", "bfloat16"])
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch between the kernels of a step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-b1", action="store_true", help="skip the batch-1 latency leg")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference GPU PyTorch path leg")
    ap.add_argument("--no-mega", action="store_true", help=argparse.SUPPRESS)      # accepted for old command lines; no effect
    ap.add_argument("--wide-pair", type=int, default=-1, help=argparse.SUPPRESS)   # A/B switch of the CTA-pair GEMM (development)
    ap.add_argument("--ref-dec-tokens", type=int, default=4, help="decode steps per bounded CPU sample")
    ap.add_argument("--no-full-report", action="store_true", help="reference arm: skip the one complete 128-token report in the warm-up")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "chat":
        run_chat(args)
    else:
        if args.wide_pair >= 0:
            from radialog_b200 import _lib
            _lib.load().rd_linear_wide_pair(args.wide_pair)
        run_own_arm(args)


if __name__ == "__main__":
    main()
