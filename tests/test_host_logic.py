"""CPU: host-side logic — weight packing (BN folding, fused layouts), sharding, and the world_size-2 data-parallel
plumbing over gloo (the N>1 path has no data-path collective: one weight broadcast at load, one gather of ids)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn.functional as F

from radialog_b200 import synth
from radialog_b200.pipeline import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 16, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_synthetic_inputs_are_deterministic_and_well_formed():
    p = synth.make_prompts(4, seed=4321)
    assert p.shape == (4, 64) and (p[:, 0] == 1).all()
    assert ((p == synth.IMG_TOKEN_ID).sum(-1) == 32).all()
    assert torch.equal(p, synth.make_prompts(4, seed=4321))
    r = synth.make_prompts(8, seed=1, ragged=True)
    assert ((r == synth.IMG_TOKEN_ID).sum(-1) == 32).all() and (r[:, 0] == 0).any()
    im = synth.make_images(2, size=64)
    assert im.shape == (2, 3, 64, 64) and torch.equal(im[:, 0], im[:, 1]) and 0 <= im.min() and im.max() < 1
    a = synth.make_llama_weights(synth.tiny_llama_cfg(), seed=0, dtype=torch.float32)
    b = synth.make_llama_weights(synth.tiny_llama_cfg(), seed=0, dtype=torch.float32)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["model.embed_tokens.weight"][0].abs().sum() == 0      # padding_idx row


def test_vision_weight_packing_folds_batchnorm_exactly():
    """pack_vision_weights: BN-folded conv == conv + eval BN; missing_previous_emb folded into the projector bias."""
    from radialog_b200.vision import pack_vision_weights, STEM_KP
    cfg = synth.tiny_vision_cfg()
    sd = synth.make_vision_weights(cfg, seed=0)
    pk = pack_vision_weights(cfg, sd, torch.float32)
    g = torch.Generator().manual_seed(0)
    R = "visual_encoder.encoder.encoder."
    x = torch.randn(2, 3, 16, 16, generator=g)
    ref = F.batch_norm(F.conv2d(x, sd[R + "conv1.weight"], stride=2, padding=3), sd[R + "bn1.running_mean"],
                       sd[R + "bn1.running_var"], sd[R + "bn1.weight"], sd[R + "bn1.bias"], False, eps=cfg.bn_eps)
    w = pk["conv1.w"][:, :147].reshape(cfg.width, 7, 7, 3).permute(0, 3, 1, 2)
    got = F.conv2d(x, w, pk["conv1.b"], stride=2, padding=3)
    assert pk["conv1.w"].shape[1] == STEM_KP and torch.allclose(got, ref, atol=1e-5)
    # projector conv1 on cat([patch, missing_previous_emb])
    E, P = "visual_encoder.encoder.", "visual_encoder.projector.model."
    patch = torch.randn(2, cfg.backbone_to_vit, 2, 2, generator=g)
    fused = torch.cat([patch, sd[E + "missing_previous_emb"].repeat(2, 1, 2, 2)], 1)
    ref = F.batch_norm(F.conv2d(fused, sd[P + "0.weight"]), sd[P + "1.running_mean"], sd[P + "1.running_var"], sd[P + "1.weight"],
                       sd[P + "1.bias"], False, eps=cfg.bn_eps)
    got = F.conv2d(patch, pk["proj1.w"][:, :, None, None], pk["proj1.b"])
    assert torch.allclose(got, ref, atol=1e-5)
    # cross-attention K/V of all cross layers concatenated in layer order
    n_cross = sum(1 for i in range(cfg.q_layers) if i % cfg.cross_attention_freq == 0)
    assert pk["q.cross_kv.w"].shape == (n_cross * 2 * cfg.q_hidden, cfg.joint_feature_size)
    # constant-folded query embedding
    B = "Qformer.bert.embeddings.LayerNorm."
    h0 = F.layer_norm(sd["query_tokens"][0], (cfg.q_hidden,), sd[B + "weight"], sd[B + "bias"], cfg.q_ln_eps)
    assert torch.allclose(pk["q.h0.w"], h0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from radialog_b200 import synth as S
    from radialog_b200.pipeline import broadcast_state_dict, gather_sequences, shard_range as sr
    from oracle import radialog_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = S.tiny_llama_cfg(num_hidden_layers=1)
    sd = S.make_llama_weights(cfg, seed=0, dtype=torch.float32) if rank == 0 else None
    sd = broadcast_state_dict(sd, src=0, device=torch.device("cpu"), bucket_bytes=1 << 20)
    ref = S.make_llama_weights(cfg, seed=0, dtype=torch.float32)
    assert set(sd) == set(ref) and all(torch.equal(sd[k], ref[k]) for k in ref), "broadcast changed the weights"
    N = 5
    prompts = S.make_prompts(N, seed=4321)
    img = torch.randn(N, 32, cfg.qformer_hidden, generator=torch.Generator().manual_seed(9))
    lo, hi = sr(N, rank, world)
    orc = O.LlamaOracle(cfg, sd, torch.float32)       # stands in for the GPU engine: the plumbing is what is under test
    local = orc.generate(prompts[lo:hi], img[lo:hi], 3, suppress_eos=True)
    counts = [sr(N, r, world)[1] - sr(N, r, world)[0] for r in range(world)]
    full = gather_sequences(local, counts, dst=0)
    if rank == 0:
        torch.save(full, os.path.join(out_dir, "gathered.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_world2_gloo(tmp_path):
    """Two ranks: weights broadcast from rank 0, units sharded, ids gathered == single-process result."""
    from oracle import radialog_oracle as O
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "gathered.pt"))
    cfg = synth.tiny_llama_cfg(num_hidden_layers=1)
    sd = synth.make_llama_weights(cfg, seed=0, dtype=torch.float32)
    prompts = synth.make_prompts(5, seed=4321)
    img = torch.randn(5, 32, cfg.qformer_hidden, generator=torch.Generator().manual_seed(9))
    want = O.LlamaOracle(cfg, sd, torch.float32).generate(prompts, img, 3, suppress_eos=True)
    assert torch.equal(got, want)


def test_embedding_pickle_and_chat_image_formats(tmp_path, monkeypatch):
    """The reference's on-disk hand-offs (pretraining/train.py:139-149, demo.py:269-273) written by the pipeline helpers are
    what modeling_llama_imgemb.py:454-462,576 reads: {dicom: np.float32[32,768]} pickles relative to the CWD, a [1,32,768] .pt."""
    import pickle

    import numpy as np
    import torch
    from radialog_b200 import pipeline
    from radialog_b200.llm import EMB_PKL_TEST, CHAT_IMG_FILE

    monkeypatch.chdir(tmp_path)
    embs = {"d1": torch.randn(32, 768, dtype=torch.float64), "d2": np.random.rand(32, 768)}
    pipeline.write_embedding_pickle(EMB_PKL_TEST, embs)
    with open(EMB_PKL_TEST, "rb") as f:
        back = pickle.load(f)
    assert sorted(back) == ["d1", "d2"]
    assert all(isinstance(v, np.ndarray) and v.dtype == np.float32 and v.shape == (32, 768) for v in back.values())
    assert np.allclose(back["d1"], embs["d1"].numpy().astype(np.float32))
    import pytest
    with pytest.raises(ValueError):
        pipeline.write_embedding_pickle(EMB_PKL_TEST, {"bad": torch.zeros(768)})
    # demo.py:269-272 saves forward_image(...)[0] = [1,32,768]; a single row is accepted too; the file is always [1,32,768]
    for src in (torch.randn(1, 32, 768), torch.randn(32, 768, dtype=torch.float16)):
        pipeline.save_chat_image(src)
        t = torch.load(CHAT_IMG_FILE)
        assert t.shape == (1, 32, 768) and t.dtype == torch.float32
        assert torch.equal(t[0], src.reshape(32, 768).float())
    with pytest.raises(ValueError):
        pipeline.save_chat_image(torch.randn(2, 32, 768))

    class FakeVision:
        def forward_image(self, images):
            b = images.shape[0]
            return images.reshape(b, -1)[:, :32 * 8].reshape(b, 32, 8), None

    batches = [(torch.arange(2 * 3 * 16 * 16, dtype=torch.float32).reshape(2, 3, 16, 16), ["a", "b"]),
               (torch.ones(1, 3, 16, 16), ["c"])]
    out = pipeline.precompute_embeddings(FakeVision(), batches, path=str(tmp_path / "x" / "emb.pkl"))
    assert sorted(out) == ["a", "b", "c"] and out["b"].shape == (32, 8)
    with open(tmp_path / "x" / "emb.pkl", "rb") as f:
        assert sorted(pickle.load(f)) == ["a", "b", "c"]
