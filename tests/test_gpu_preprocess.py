"""GPU parity of the device-side image preprocessing (csrc/preprocess.cu through the C-ABI) against the numpy oracle and the
fixtures made with the reference's own Pillow + torchvision stack.  Byte / integer work: the bar is bit-exact."""
import glob
import os

import numpy as np
import pytest
import torch

from radialog_b200.preprocess import ChestXrayTransform, create_chest_xray_transform_for_inference
from oracle import preprocess_oracle as P
from oracle.make_golden_preprocess import synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tf(cuda_dev):
    return create_chest_xray_transform_for_inference(512, center_crop_size=448, device=cuda_dev)


def test_reference_fixtures_bit_exact(tf, golden_dir):
    for f in sorted(glob.glob(os.path.join(golden_dir, "preprocess_*.npz"))):
        z = np.load(f)
        a = synth_image(int(z["h"]), int(z["w"]), np.dtype(str(z["dtype"])).type, int(z["seed"]))
        out = tf(a).cpu().numpy()
        want = z["plane_u8"].astype(np.float32) / np.float32(255.0)
        assert out.shape == (3, 448, 448)
        for c in range(3):
            assert np.array_equal(out[c], want), os.path.basename(f)


@pytest.mark.parametrize("h,w,dt", [(512, 512, np.uint8), (448, 600, np.uint16), (600, 448, np.float32), (1500, 1201, np.uint16),
                                    (513, 2048, np.uint8), (300, 333, np.uint16), (3056, 2544, np.uint16)])
def test_random_images_vs_oracle(tf, h, w, dt):
    """Seeded noise images (worst case for resampling), portrait / landscape / square / upscaling / full-size radiograph."""
    rng = np.random.default_rng(h * 31 + w)
    a = rng.standard_normal((h, w)).astype(np.float32) if dt == np.float32 else rng.integers(0, np.iinfo(dt).max, size=(h, w)).astype(dt)
    out = tf(a)
    want = P.preprocess(a)
    assert np.array_equal(out.cpu().numpy(), want)
    assert tf.launch_count() > 0


def test_already_remapped_pil_style_input_and_errors(tf):
    rng = np.random.default_rng(5)
    u8 = rng.integers(0, 256, size=(640, 700)).astype(np.uint8)
    nh, nw = P.resized_size(640, 700, 512)
    r = P.pil_resize_bilinear(u8, nh, nw)
    top, left = P.center_crop_offsets(nh, nw, 448)
    want = r[top:top + 448, left:left + 448].astype(np.float32) / np.float32(255.0)
    out = tf(torch.from_numpy(u8), remap=False).cpu().numpy()
    assert np.array_equal(out[0], want) and np.array_equal(out[1], want) and np.array_equal(out[2], want)
    with pytest.raises(ValueError):
        tf(np.zeros((3, 10, 10), np.uint8))                 # ExpandChannels-style shape check (ReportDataset.py:92-93)
    with pytest.raises(ValueError):
        tf(np.zeros((10, 10), np.int32))
    with pytest.raises(RuntimeError):
        tf(np.zeros((5000, 10), np.uint8))                  # larger than the handle was created for


def test_feeds_forward_image(cuda_dev):
    """The tensor goes straight into forward_image (demo.py:269): same result as handing over the oracle's tensor."""
    from radialog_b200 import synth
    from radialog_b200.vision import Blip2Qformer
    vcfg = synth.tiny_vision_cfg(q_hidden=768, q_heads=12, q_intermediate=256, joint_feature_size=128)
    vis = Blip2Qformer.from_state_dict(vcfg, synth.make_vision_weights(vcfg, seed=0), device=cuda_dev, max_batch=1)
    s = vcfg.image_size
    tf_small = ChestXrayTransform(resize=s + s // 8, center_crop_size=s, device=cuda_dev)
    a = synth_image(700, 560, np.uint16, 3)
    want = P.preprocess(a, s + s // 8, s)
    got = tf_small(a)
    assert np.array_equal(got.cpu().numpy(), want)
    q_gpu, _ = vis.forward_image(got[None])
    q_ref, _ = vis.forward_image(torch.from_numpy(want)[None].to(cuda_dev))
    assert torch.equal(q_gpu, q_ref)


def test_throughput_is_reported(tf, cuda_dev, capsys):
    """Full-size radiograph (3056 x 2544 uint16, 15.5 MB): device time per image with the raw image already in HBM."""
    rng = np.random.default_rng(1)
    a = torch.from_numpy(rng.integers(0, 4096, size=(3056, 2544)).astype(np.int16)).view(torch.uint16).to(cuda_dev)
    out = torch.empty(3, 448, 448, device=cuda_dev)
    for _ in range(3):
        tf(a, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        tf(a, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    with capsys.disabled():
        print(f"\n[preprocess] 3056x2544 uint16 -> [3,448,448] fp32: {us:.0f} us per image ({a.numel() * 2 / us / 1e3:.0f} GB/s of raw image read)")
    assert us < 5000
