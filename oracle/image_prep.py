"""CPU restatement (numpy, float32) of the reference's centre-crop step of image preparation.

TEST INFRASTRUCTURE ONLY (the checker of the device-side crop kernel; never on the product path).

Reference: experiments/robot/openvla_utils.py
  center_crop_image   :616-648   uint8 -> float32 in [0,1] (tf.image.convert_image_dtype), crop_and_resize with
                                 crop_scale = 0.9, clip to [0,1], back to uint8 (convert_image_dtype, saturate=True)
  crop_and_resize     :568-613   box = centred square of side sqrt(crop_scale) (in normalised coordinates), resampled
                                 to 224 x 224 by tf.image.crop_and_resize (bilinear)
The arithmetic lives in TensorFlow (un-vendored dependency `tensorflow==2.15.0`, pyproject.toml) which is absent from
this image: it is restated from TF's published kernel semantics -
  convert_image_dtype(uint8 -> float32): x * (1 / 255)
  crop_and_resize (CPU kernel, method="bilinear", extrapolation unused for a centred box):
      scale   = (y2 - y1) * (H - 1) / (crop_h - 1)
      in_y    = y1 * (H - 1) + y * scale;  top = floor(in_y), bottom = ceil(in_y), lerp = in_y - top   (same in x)
      value   = top_row + (bottom_row - top_row) * y_lerp,   row = left + (right - left) * x_lerp
  convert_image_dtype(float32 -> uint8, saturate=True): saturate_cast(x * 255.5)  (truncation after scaling by max + 0.5)
PARITY UNPINNED: no TensorFlow here to generate a golden image; the restatement is checked by its own properties and
against an independent implementation of the same sampling rule (torch grid_sample, align_corners=True: equal to one
uint8 step, > 99.5 % of pixels exactly; tests/test_image_prep.py), and the device kernel is bit-exact against it.  After this step the image is already
224 x 224, so the processor's resize / centre-crop (processing_prismatic.py:136-137) are identities and ToTensor +
Normalize follow (the engine's lookup table, include/vla_b200.h vla_predict_u8).
The first half of the reference's preparation - JPEG encode / decode and the lanczos3 antialiased resize from the camera
resolution (:560-565) - depends on TF's JPEG codec and stays on the CPU.
"""
from __future__ import annotations

import numpy as np

OPENVLA_IMAGE_SIZE = 224


def crop_box(crop_scale: float = 0.9):
    """(y1, x1, y2, x2) of openvla_utils.py:588-603 in float32 arithmetic."""
    side = np.clip(np.sqrt(np.float32(crop_scale)), np.float32(0), np.float32(1)).astype(np.float32)
    off = ((np.float32(1) - side) / np.float32(2)).astype(np.float32)
    return off, off, (off + side).astype(np.float32), (off + side).astype(np.float32)


def center_crop_image(img: np.ndarray, crop_scale: float = 0.9, out_size: int = OPENVLA_IMAGE_SIZE) -> np.ndarray:
    """img: (H, W, 3) uint8 -> (out_size, out_size, 3) uint8, openvla_utils.py:616-648."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    x = img.astype(np.float32) * np.float32(1.0 / 255.0)
    y1, x1, y2, x2 = crop_box(crop_scale)
    f32 = np.float32
    hs = ((y2 - y1) * f32(H - 1) / f32(out_size - 1)).astype(np.float32) if out_size > 1 else f32(0)
    ws = ((x2 - x1) * f32(W - 1) / f32(out_size - 1)).astype(np.float32) if out_size > 1 else f32(0)
    ii = np.arange(out_size, dtype=np.float32)
    in_y = (y1 * f32(H - 1) + ii * hs).astype(np.float32)
    in_x = (x1 * f32(W - 1) + ii * ws).astype(np.float32)
    top, bot = np.floor(in_y).astype(np.int64), np.ceil(in_y).astype(np.int64)
    left, right = np.floor(in_x).astype(np.int64), np.ceil(in_x).astype(np.int64)
    yl = (in_y - top.astype(np.float32)).astype(np.float32)[:, None, None]
    xl = (in_x - left.astype(np.float32)).astype(np.float32)[None, :, None]
    tl, tr = x[top][:, left], x[top][:, right]
    bl, br = x[bot][:, left], x[bot][:, right]
    t = (tl + (tr - tl) * xl).astype(np.float32)
    b = (bl + (br - bl) * xl).astype(np.float32)
    v = (t + (b - t) * yl).astype(np.float32)
    v = np.clip(v, f32(0), f32(1))
    return np.clip(np.trunc(v * f32(255.5)), 0, 255).astype(np.uint8)
