"""Image -> report in one call, plus the data-parallel plumbing.

``ReportPipeline.generate`` is demo.py:269-297 / test.py:336-348 with the Q-Former tokens handed over as a device tensor
instead of the reference's pickle / ``current_chat_img.pt`` round trip.  Multi-GPU is plain data parallelism over
independent (image, prompt) units: ``shard_range`` splits the batch, ``broadcast_state_dict`` is the single collective
(weights at load, NCCL on GPUs / gloo in the CPU tests) and ``gather_sequences`` collects the token ids at the end; the
decode step itself has no collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import os

import torch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of ``n`` units for ``rank`` (first ``n % world`` ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], src: int = 0, device: Optional[torch.device] = None,
                         bucket_bytes: int = 1 << 30) -> Dict[str, torch.Tensor]:
    """Rank ``src`` holds ``sd``; every rank returns an identical copy.  Tensors are packed into ~1 GiB byte buckets so
    the 13.5 GB of fp16 weights go out in a handful of broadcasts (NVLink/NVSwitch: launch latency, not link count,
    is what the bucket size trades against)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert sd is not None
        return sd
    rank = dist.get_rank()
    meta = [[(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()]] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    out: Dict[str, torch.Tensor] = {}
    bucket: List[Tuple[str, Tuple[int, ...], torch.dtype, int]] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.empty(size, dtype=torch.uint8, device=device)
        if rank == src:
            off = 0
            for k, shape, dt, nbytes in bucket:
                flat[off:off + nbytes] = sd[k].to(device).contiguous().view(-1).view(torch.uint8)
                off += (nbytes + 15) // 16 * 16
        dist.broadcast(flat, src=src)
        off = 0
        for k, shape, dt, nbytes in bucket:
            out[k] = flat[off:off + nbytes].view(dt).view(shape).clone() if rank != src else sd[k]
            off += (nbytes + 15) // 16 * 16
        bucket, size = [], 0

    for k, shape, dtname in meta[0]:
        dt = getattr(torch, dtname)
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dt).element_size()
        padded = (nbytes + 15) // 16 * 16
        if size + padded > bucket_bytes and bucket:
            flush()
        bucket.append((k, shape, dt, nbytes))
        size += padded
    flush()
    return out


def gather_sequences(local: torch.Tensor, counts: List[int], dst: int = 0) -> Optional[torch.Tensor]:
    """Collects per-rank ``int64[B_local, L]`` token ids on ``dst`` in rank order (rows may differ per rank, L must not)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    L = local.shape[1]
    mx = max(counts)
    pad = torch.zeros(mx, L, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def write_embedding_pickle(path: str, embeddings: Dict[str, "torch.Tensor"]) -> None:
    """The reference's on-disk hand-off of Q-Former outputs (pretraining/train.py:139-149, read back by
    modeling_llama_imgemb.py:454-462): a pickled ``{dicom_id: np.float32[32, 768]}`` dict."""
    import pickle

    import numpy as np
    out = {}
    for k, v in embeddings.items():
        a = v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float32)
        if a.ndim != 2:
            raise ValueError(f"embedding of {k!r} must be [num_query_tokens, hidden], found {a.shape}")
        out[str(k)] = np.ascontiguousarray(a, dtype=np.float32)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(out, f)


@torch.no_grad()
def precompute_embeddings(vision, batches, path: Optional[str] = None) -> Dict[str, "torch.Tensor"]:
    """Bulk precompute of the Q-Former outputs - the evaluate branch of pretraining/train.py:134-173: ``batches`` yields
    ``(images [B,3,448,448], dicom ids)``; returns ``{dicom: float32 [32,768] (cpu)}`` and, with ``path``, writes the pickle
    in the reference format (e.g. ``pretraining/embs/<run>_embeddings_test.pkl``)."""
    embs: Dict[str, torch.Tensor] = {}
    for images, ids in batches:
        q_out, _ = vision.forward_image(images)
        q_cpu = q_out.float().cpu()
        for j, d in enumerate(ids):
            embs[str(d)] = q_cpu[j].clone()
    if path is not None:
        write_embedding_pickle(path, embs)
    return embs


def save_chat_image(q_out: "torch.Tensor", path: str = "current_chat_img.pt") -> None:
    """demo.py:269-272: the conversational path saves ``forward_image(image)[0]`` - a ``[1, 32, 768]`` float32 tensor - to
    ``current_chat_img.pt`` in the CWD; ``generate(use_img=True)`` reads it back (modeling_llama_imgemb.py:576), and the
    reference's reader iterates dim 0 as the batch.  Accepts ``[1, 32, 768]`` (the reference call, verbatim) or one row
    ``[32, 768]``; always writes ``[1, 32, 768]`` float32 so that the file also feeds the reference's own reader."""
    if q_out.dim() == 2:
        q_out = q_out[None]
    if q_out.dim() != 3 or q_out.shape[0] != 1:
        raise ValueError(f"expected [1, num_query_tokens, hidden] or [num_query_tokens, hidden], found {tuple(q_out.shape)}")
    torch.save(q_out.detach().float().cpu().contiguous(), path)


class ReportPipeline:
    def __init__(self, vision, llm):
        self.vision = vision
        self.llm = llm
        self.last_stats: Dict[str, float] = {}

    @torch.no_grad()
    def generate(self, images: torch.Tensor, input_ids: torch.Tensor, max_new_tokens: int = 128, suppress_eos: bool = False,
                 return_dict_in_generate: bool = False, output_scores: bool = False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        q_out, _ = self.vision.forward_image(images)
        ev[1].record()
        out = self.llm.generate(input_ids, img_embeds=q_out, max_new_tokens=max_new_tokens, suppress_eos=suppress_eos,
                                return_dict_in_generate=return_dict_in_generate, output_scores=output_scores)
        torch.cuda.synchronize()
        self.last_stats = dict(self.llm.last_stats, vision_ms=ev[0].elapsed_time(ev[1]))
        return out
