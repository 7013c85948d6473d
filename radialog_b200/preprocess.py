"""Device-side replacement of the reference's image preprocessing.

Reference: ``demo.py:206-218`` (``load_image``: ``remap_to_uint8`` -> PIL "L") followed by
``create_chest_xray_transform_for_inference(512, center_crop_size=448)`` (``model/lavis/data/ReportDataset.py:97-106``:
``Resize`` -> ``CenterCrop`` -> ``ToTensor`` -> ``ExpandChannels``).  Here the raw grey image is copied to the GPU once and
``rd_preproc_run`` produces the float32 ``[3, 448, 448]`` tensor ``forward_image`` takes, bit-identical to the CPU pipeline.
There is no CPU fallback: without the CUDA library the constructor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Union

import numpy as np
import torch

from . import _lib

_DT = {torch.uint8: 0, torch.uint16: 1, torch.float32: 2}


class ChestXrayTransform:
    """Callable with the reference transform's signature: image in, ``[3, crop, crop]`` float32 tensor out (on the GPU)."""

    def __init__(self, resize: int = 512, center_crop_size: int = 448, device: Union[str, torch.device] = "cuda:0",
                 max_size: int = 4096):
        if not torch.cuda.is_available():
            raise RuntimeError("radialog_b200.preprocess needs a CUDA device (no CPU fallback)")
        self.resize, self.crop, self.device = int(resize), int(center_crop_size), torch.device(device)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rd_preproc_create(max_size, max_size, self.resize, self.crop, C.byref(self._h)), "preproc_create")

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._lib.rd_preproc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self._lib.rd_preproc_launch_count(self._h))

    @torch.no_grad()
    def __call__(self, image, remap: bool = None, out: torch.Tensor = None) -> torch.Tensor:
        """image: ``[H, W]`` grey image - numpy array / torch tensor (uint8, uint16 or float32), or a PIL "L" image.
        ``remap`` (default: True for arrays = the raw file content as ``io.imread`` returns it, False for PIL images, which
        ``load_image`` has already remapped) applies ``remap_to_uint8`` first."""
        if hasattr(image, "mode") and hasattr(image, "size") and not isinstance(image, (np.ndarray, torch.Tensor)):   # PIL image
            if image.mode != "L":
                raise ValueError(f"expected a mode 'L' image, got {image.mode!r}")
            image = np.asarray(image)
            remap = False if remap is None else remap
        remap = True if remap is None else remap
        if isinstance(image, np.ndarray):
            if image.dtype == np.float64:
                image = image.astype(np.float32)
            if image.dtype == np.uint16:
                t = torch.from_numpy(image.view(np.int16)).view(torch.uint16)
            else:
                t = torch.from_numpy(np.ascontiguousarray(image))
        else:
            t = image
        if t.dim() != 2:
            raise ValueError(f"Expected a grey image of shape [H, W], found {tuple(t.shape)}")
        if t.dtype not in _DT:
            raise ValueError(f"unsupported image dtype {t.dtype} (uint8, uint16 or float32)")
        if not remap and t.dtype != torch.uint8:
            raise ValueError("remap=False needs a uint8 image")
        t = t.to(self.device).contiguous()
        H, W = int(t.shape[0]), int(t.shape[1])
        if out is None:
            out = torch.empty(3, self.crop, self.crop, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rd_preproc_run(self._h, t.data_ptr(), _DT[t.dtype], H, W, 1 if remap else 0, out.data_ptr(),
                                                _lib.current_stream()), "preproc_run")
        return out


def create_chest_xray_transform_for_inference(resize: int, center_crop_size: int, device="cuda:0") -> ChestXrayTransform:
    """Same name and arguments as ``model/lavis/data/ReportDataset.py:97``."""
    return ChestXrayTransform(resize, center_crop_size, device)
