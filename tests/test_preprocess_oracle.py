"""CPU: the numpy oracle of the image preprocessing (oracle/preprocess_oracle.py) against fixtures made with the reference's
own transform stack (Pillow + torchvision; oracle/make_golden_preprocess.py) and, where Pillow is importable, against it live."""
import glob
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as P
from oracle.make_golden_preprocess import synth_image


def test_oracle_matches_reference_fixtures(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "preprocess_*.npz")))
    assert len(files) >= 3
    for f in files:
        z = np.load(f)
        a = synth_image(int(z["h"]), int(z["w"]), np.dtype(str(z["dtype"])).type, int(z["seed"]))
        out = P.preprocess(a)
        assert out.shape == (3, 448, 448) and out.dtype == np.float32
        want = z["plane_u8"].astype(np.float32) / np.float32(255.0)
        for c in range(3):
            assert np.array_equal(out[c], want), f"{os.path.basename(f)}: channel {c} differs from the reference pipeline"


@pytest.mark.parametrize("h,w", [(512, 512), (448, 600), (1500, 1201), (513, 2048)])
def test_oracle_matches_pillow_live(h, w):
    PIL = pytest.importorskip("PIL.Image")
    tv = pytest.importorskip("torchvision.transforms")
    rng = np.random.default_rng(h * 7 + w)
    a = rng.integers(0, 4096, size=(h, w)).astype(np.uint16)
    u8 = P.remap_to_uint8(a)
    ref = tv.Compose([tv.Resize(512), tv.CenterCrop(448), tv.ToTensor()])(PIL.fromarray(u8).convert("L")).numpy()[0]
    assert np.array_equal(P.preprocess(a)[0], ref)


def test_remap_and_geometry_details():
    a = np.array([[10, 20], [30, 110]], np.uint16)
    assert P.remap_to_uint8(a).tolist() == [[0, 25], [51, 255]]          # truncation, not rounding (demo.py:203)
    assert P.resized_size(700, 560, 512) == (640, 512) and P.resized_size(520, 800, 512) == (512, 787)
    assert P.center_crop_offsets(787, 512, 448) == (170, 32)               # round-half-to-even: (787-448)/2 = 169.5 -> 170
    assert P.center_crop_offsets(785, 512, 448) == (168, 32)               # 168.5 -> 168
