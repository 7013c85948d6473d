#!/usr/bin/env python
"""Per-phase timeline of the persistent decode kernel (decode_mega.cu) at Vicuna-7B size (development tool).

    python tools/trace_mega.py [B] [extra_ctx]

Every CTA stamps %globaltimer at fixed points of each (layer, phase); this prints, averaged over layers 1.. and CTAs:
how long a phase lasts for the whole grid, how long the grid barrier takes after the last arrival, when the first MMA
is issued and the last one committed, and how long the epilogues take."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radialog_b200 import _lib, synth  # noqa: E402
from radialog_b200.llm import LlamaForCausalLM  # noqa: E402

dev = torch.device("cuda:0")
dtype = torch.bfloat16
lib = _lib.load()
lib.rd_mega_set_trace.argtypes = [C.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 6
lib.rd_set_pdl(int(os.environ.get("PDL", "1")))
cfg = synth.LlamaCfg()
sd = synth.make_llama_weights(cfg, seed=0, dtype=dtype, device="cuda:0")
llm = LlamaForCausalLM.from_state_dict(cfg, sd, torch_dtype=dtype, device=dev)
del sd
llm.use_cuda_graph = False
llm.set_mega(True)
prompts = synth.make_prompts(B, seed=4321).to(dev)
img = torch.randn(B, 32, 768, device=dev) * 0.5
llm.reserve(B, 64 + extra + 8)
llm.generate(prompts, img_embeds=img, max_new_tokens=extra, suppress_eos=True)
torch.cuda.synchronize()
G, L = 148, cfg.num_hidden_layers
trace = torch.zeros(G * L * 5 * 16 + 320 + G * 12 + G * 64, dtype=torch.int64, device=dev)
st = _lib.current_stream()
_lib.check(lib.rd_llm_decode_step(llm._h, st), "decode_step")
torch.cuda.synchronize()
lib.rd_mega_set_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(lib.rd_llm_decode_step(llm._h, st), "decode_step")
e1.record()
torch.cuda.synchronize()
lib.rd_mega_set_trace(None)
t = trace[:G * L * 5 * 16].view(G, L, 5, 16).cpu().double()
kbt = trace[G * L * 5 * 16:].cpu().double()
print(f"B={B} ctx~{64 + extra}: traced decode step {e0.elapsed_time(e1) * 1e3:.0f} us (all kernels of the step)")
names = ["qkv", "attn", "o", "gate_up", "down"]


def stat(x):
    x = x[x > 0]
    return (x.min().item(), x.mean().item(), x.max().item()) if x.numel() else (0, 0, 0)


layers = range(1, L)
tot = {}
for ph in range(5):
    rows = []
    for l in layers:
        ev = t[:, l, ph, :]                                    # [G, 8]
        start = ev[:, 0]
        first_start = start[start > 0].min().item()
        last_start = start.max().item()
        arrive = ev[:, 3]
        last_arrive = arrive.max().item()
        first_arrive = arrive[arrive > 0].min().item()
        # previous phase's last arrival
        pl, pp = (l, ph - 1) if ph > 0 else (l - 1, 4)
        prev_last = t[:, pl, pp, 3].max().item()
        d = {"span": last_arrive - prev_last, "bar_lat": first_start - prev_last, "start_skew": last_start - first_start,
             "arrive_skew": last_arrive - first_arrive}
        if ph != 1:
            work = ev[:, 4] > 0
            d["first_mma"] = (ev[work, 4] - prev_last).mean().item()
            d["last_mma"] = (ev[work, 5] - prev_last).mean().item()
            d["last_mma_max"] = (ev[work, 5] - prev_last).max().item()
            d["xprod"] = (ev[work, 6] - prev_last).mean().item()
            d["xform_done"] = (ev[work, 1] - prev_last).mean().item()
            d["acc_ready"] = (ev[work, 2] - prev_last).mean().item()
            d["epi"] = (ev[work, 3] - ev[work, 2]).mean().item()
            d["w_first_issue"] = (ev[work, 7] - prev_last).mean().item()
            fin = ev[:, 8] > 0
            if fin.any():
                slow = ev[fin, 3].argmax()
                e = ev[fin][slow]
                # the finalising CTA that arrives last: last MMA commit -> accumulator seen -> contributors seen -> partials summed -> arrive issued
                d["L:mma"] = e[5].item() - prev_last
                d["L:ctr"] = e[8].item() - prev_last
                d["L:sum"] = e[9].item() - prev_last
                d["L:pre_arrive"] = e[3].item() - prev_last
                d["L:arrived"] = e[10].item() - prev_last
        rows.append(d)
    keys = rows[0].keys()
    avg = {k: sum(r[k] for r in rows) / len(rows) / 1e3 for k in keys}
    tot[names[ph]] = avg
    print(f"{names[ph]:8s} " + "  ".join(f"{k} {v:6.1f}" for k, v in avg.items()))
print("sum of phase spans per layer: %.1f us" % sum(v["span"] for v in tot.values()))

# per-k-block stamps of CTA 0, layer 1, QKV phase: weight producer issue times and MMA "operands ready" times
iss, rdy = kbt[:64], kbt[64:128]
n = int((rdy > 0).sum().item())
if n > 2:
    base = iss[0].item()
    print("CTA0 layer1 qkv: k-block i: W issue time / MMA ready time (us since first issue)")
    print("  issue: " + " ".join(f"{(iss[i].item() - base) / 1e3:5.1f}" for i in range(n)))
    print("  ready: " + " ".join(f"{(rdy[i].item() - base) / 1e3:5.1f}" for i in range(n)))
    for nm, off in (("loop top", 256), ("mma issued", 128), ("committed", 192)):
        print(f"  {nm}: " + " ".join(f"{(kbt[off + i].item() - base) / 1e3:5.1f}" for i in range(n)))

cyc = kbt[320:320 + G * 8].view(G, 8)
m = cyc.mean(0)
nk = max(1.0, m[5].item())
print("MMA thread, mean over CTAs (cycles per k-block): wait W %.0f  wait X %.0f  fence+MMA issue %.0f  commit %.0f | acc_empty wait total %.0f  k-blocks %.0f  thread total %.0f cycles"
      % (m[0] / nk, m[1] / nk, m[2] / nk, m[3] / nk, m[4], nk, m[6]))

spw = (trace[G * L * 5 * 16 + 320:G * L * 5 * 16 + 320 + G * 8].view(G, 8)[:, 7].cpu() >> 32).double().mean().item()
spx = (trace[G * L * 5 * 16 + 320:G * L * 5 * 16 + 320 + G * 8].view(G, 8)[:, 7].cpu() & 0xFFFFFFFF).double().mean().item()
print("MMA thread try_wait probes per k-block: W %.2f  X %.2f" % (spw / nk, spx / nk))
wp = kbt[320 + G * 8:320 + G * 12].view(G, 4).mean(0)
print("W producer: tiles %.0f, cycles waiting for a free slot per tile %.0f (probes %.2f), thread total %.0f cycles" %
      (wp[3], wp[0] / max(1, wp[3]), wp[2] / max(1, wp[3]), wp[1]))

att = kbt[320 + G * 12:320 + G * 12 + G * 64].view(G * 8, 8)
busy = att[att[:, 0] > 0]
if busy.numel():
    mm = busy.mean(0) / L / 1.965e3
    print("attention per warp with work, us per layer: prologue(RoPE/LoRA/append) %.1f  scores %.1f  max+sum %.1f  P.V %.1f  combine/write %.1f   (%d of %d warps busy)"
          % (mm[0], mm[1], mm[2], mm[3], mm[4], busy.shape[0], G * 8))
