#!/usr/bin/env python
"""Generates tests/golden/preprocess_*.npz with the reference's OWN transform stack - Pillow + torchvision, exactly as
demo.py:206-218 + ReportDataset.py:97-106 call them - on seeded synthetic grey images, so that the numpy oracle
(oracle/preprocess_oracle.py) and the CUDA path are pinned to it.  Run in the build container:  python -m oracle.make_golden_preprocess"""
import os

import numpy as np
import torch
from PIL import Image
from torchvision.transforms import CenterCrop, Compose, Resize, ToTensor

from oracle import preprocess_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("u16_portrait", 700, 560, np.uint16, 11), ("u8_landscape", 520, 800, np.uint8, 12), ("f32_upscale", 300, 420, np.float32, 13)]


def synth_image(h, w, dtype, seed):
    """Smooth chest-X-ray-like content (low-frequency blobs + mild noise) so the fixture compresses well."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w))
    for _ in range(6):
        cy, cx, s = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(40, 200)
        img += rng.uniform(0.2, 1.0) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    img += 0.02 * rng.standard_normal((h, w))
    img = (img - img.min()) / (img.max() - img.min())
    if dtype == np.float32:
        return (img * 3.0 - 1.0).astype(np.float32)
    return (img * (np.iinfo(dtype).max * 0.9)).astype(dtype)


def reference_pipeline(array):
    u8 = P.remap_to_uint8(array)                                    # demo.py:217 (the function itself is restated, 4 lines)
    pil = Image.fromarray(u8).convert("L")                          # demo.py:218
    t = Compose([Resize(512), CenterCrop(448), ToTensor()])(pil)    # ReportDataset.py:104
    return torch.repeat_interleave(t, 3, dim=0).numpy()             # ExpandChannels, ReportDataset.py:94


def main():
    for name, h, w, dt, seed in CASES:
        a = synth_image(h, w, dt, seed)
        out = reference_pipeline(a)
        plane = np.rint(out[0] * 255.0).astype(np.uint8)
        assert np.array_equal(plane.astype(np.float32) / np.float32(255.0), out[0]) and np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])
        path = os.path.join(ROOT, "tests", "golden", f"preprocess_{name}.npz")
        np.savez_compressed(path, h=h, w=w, dtype=np.dtype(dt).name, seed=seed, plane_u8=plane)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
