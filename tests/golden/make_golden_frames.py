"""Generates tests/golden/frame_prep.npz: outputs of OpenCV's own cv2.sepFilter2D / cv2.resize (the arithmetic behind
style_transfer.py:124-130) on small seeded uint8 frames, used to pin oracle.sep_filter_1331_u8 / resize_linear_u8 bit-exactly.
Run in the build container (cv2 is not needed on the GPU box):  python tests/golden/make_golden_frames.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
kernel_1d = np.array([[0.125], [0.375], [0.375], [0.125]])      # style_transfer.py:97

rng = np.random.default_rng(20260923)
out = {"cv2_version": np.array(cv2.__version__)}
cases = [(37, 53, 0.7, (30, 22), (2, 20, 3, 28)),     # H, W, scale, (w, h), (top, bottom, left, right)
         (64, 48, 0.3, (17, 23), (0, 23, 0, 17)),
         (25, 31, 1.4, (44, 36), (5, 30, 4, 40)),      # up-scaling (small faces): no blur
         (90, 120, 0.5, (60, 45), (1, 44, 2, 59)),
         (3, 5, 0.2, (2, 2), (0, 2, 0, 2))]
for i, (H, W, scale, size, crop) in enumerate(cases):
    f = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if i == 3:
        f[::2] = 255 - (f[::2] // 64)            # saturating / tie-heavy content
    g = f
    if scale <= 0.75:
        g = cv2.sepFilter2D(g, -1, kernel_1d, kernel_1d)
        out[f"c{i}_blur1"] = g
    if scale <= 0.375:
        g = cv2.sepFilter2D(g, -1, kernel_1d, kernel_1d)
        out[f"c{i}_blur2"] = g
    r = cv2.resize(g, size)
    top, bottom, left, right = crop
    out[f"c{i}_frame"] = f
    out[f"c{i}_params"] = np.array([scale, size[0], size[1], top, bottom, left, right], dtype=np.float64)
    out[f"c{i}_resized"] = r
    out[f"c{i}_out"] = r[top:bottom, left:right]
out["n_cases"] = np.array(len(cases))
np.savez_compressed(os.path.join(HERE, "frame_prep.npz"), **out)
print("wrote frame_prep.npz", {k: getattr(v, "shape", None) for k, v in out.items() if k.endswith("_out")})
