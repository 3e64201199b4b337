"""CPU restatement of `resize_by_factor` (empanada/data/utils/transforms.py:9-21):
`cv2.resize(image, (ceil(w/f), ceil(h/f)))` with OpenCV's default INTER_LINEAR.
TEST INFRASTRUCTURE ONLY (see oracle/post.py header).

PARITY UNPINNED: OpenCV (`opencv-python`, unpinned in the reference's setup.cfg) is a third-party
dependency that is neither vendored in /root/reference nor installed in this image, and no
reference test holds a golden vector for it. This restates the published 8-bit algorithm of
OpenCV's imgproc/resize.cpp (non-IPP build): pixel-centre mapping `src = (dst + 0.5) * scale -
0.5`, 11-bit fixed-point coefficients rounded half-to-even, horizontal pass in int32, vertical
pass `((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2`; an exact 2 x 2 reduction takes the
INTER_AREA fast path `(a+b+c+d+2) >> 2`, to which OpenCV switches INTER_LINEAR.
"""
import math

import numpy as np


def _coef(n_dst, n_src, clamp_edges):
    scale = 1.0 / (float(n_dst) / float(n_src))
    f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_edges:
        lo = s < 0
        f[lo], s[lo] = 0.0, 0
        hi = s >= n_src - 1
        f[hi], s[hi] = 0.0, n_src - 1
    c0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int64)
    c1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    return s, c0, c1


def resize_linear_u8(image, dh, dw):
    h, w = image.shape
    img = image.astype(np.int64)
    if h == 2 * dh and w == 2 * dw:
        return ((img[0::2, 0::2] + img[0::2, 1::2] + img[1::2, 0::2] + img[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = _coef(dw, w, True)
    sy, b0, b1 = _coef(dh, h, False)
    x1 = np.minimum(sx + 1, w - 1)
    y0 = np.clip(sy, 0, h - 1)
    y1 = np.clip(sy + 1, 0, h - 1)
    r0 = img[y0][:, sx] * a0[None, :] + img[y0][:, x1] * a1[None, :]
    r1 = img[y1][:, sx] * a0[None, :] + img[y1][:, x1] * a1[None, :]
    v = (((b0[:, None] * (r0 >> 4)) >> 16) + ((b1[:, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def resize_by_factor(image, scale_factor=1):
    """data/utils/transforms.py:9-21."""
    if scale_factor == 1:
        return image
    if image.dtype != np.uint8:
        raise NotImplementedError("only OpenCV's 8-bit path is restated")
    h, w = image.shape
    return resize_linear_u8(image, math.ceil(h / scale_factor), math.ceil(w / scale_factor))
