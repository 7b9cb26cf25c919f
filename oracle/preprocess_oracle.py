"""TEST INFRASTRUCTURE — CPU restatement of the reference's image pre-processing.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module; the product path
(`emu_preprocess_image` in emu_b200/csrc/preprocess.cu) never does.

Reference call sites: `TF.Resize((448, 448), interpolation=BICUBIC) -> TF.ToTensor() -> TF.Normalize(mean, std)` at
Emu2/emu/chat.py:35-39, Emu2/emu/diffusion.py:59-63, Emu1/models/pipeline.py:59-63 (224).  On a PIL image torchvision's
Resize is `PIL.Image.resize(size, BICUBIC)` (third-party: Pillow, unpinned in Emu2/requirements.txt:5; 12.2.0 installed
here), i.e. libImaging/Resample.c `ImagingResample`: a separable two-pass convolution in fixed point on uint8,
horizontal pass first, every pass rounded and clipped to uint8.  This file restates that algorithm:

  * precompute_coeffs   (Resample.c `precompute_coeffs`): per output pixel, window [xmin, xmin+xmax) and double weights
    w = bicubic((x + xmin - center + 0.5) / filterscale), normalised by their sum; support = 2 * max(scale, 1)
  * normalize_coeffs_8bpc: k_int = (int)(±0.5 + k * 2^22)  (PRECISION_BITS = 32 - 8 - 2)
  * resample pass: out = clip8((2^21 + sum_x in[x] * k_int[x]) >> 22)
  * ToTensor: uint8 -> float32 / 255 ; Normalize: (x - mean) / std in float32 (torchvision functional order)

PINNED: tests/test_oracle_cpu.py::test_preprocess_oracle_vs_torchvision runs the reference's own transform
(torchvision + Pillow, both present in the authoring container) on random images and demands bit-exact uint8 after the
resize and bit-exact float32 after the normalisation.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """-> ksize, bounds [out,2] (xmin, count), integer coefficients [out, ksize] (int32), as Pillow computes them for a
    full-image box (in0 = 0, in1 = in_size)."""
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img [H, W, C] uint8 -> [out_h, out_w, C] uint8, bit-exact with PIL.Image.resize((out_w, out_h), BICUBIC)."""
    H, W, C = img.shape
    x = img.astype(np.int64)
    if out_w != W:
        _, bx, kx = precompute_coeffs(W, out_w)
        tmp = np.empty((H, out_w, C), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = int(bx[xx, 0]), int(bx[xx, 1])
            acc = (x[:, x0:x0 + n, :] * kx[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        x = tmp.astype(np.int64)
    if out_h != H:
        _, by, ky = precompute_coeffs(H, out_h)
        out = np.empty((out_h, x.shape[1], C), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = int(by[yy, 0]), int(by[yy, 1])
            acc = (x[y0:y0 + n, :, :] * ky[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        return out
    return x.astype(np.uint8)


def to_tensor_normalize(img_u8: np.ndarray, mean, std) -> np.ndarray:
    """[H, W, 3] uint8 -> [3, H, W] float32: ToTensor (/255 in fp32) then Normalize ((x - mean) / std in fp32)."""
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return ((x - m) / s).astype(np.float32)


def image_transform(img_u8: np.ndarray, size: int, mean, std) -> np.ndarray:
    return to_tensor_normalize(resize_bicubic_u8(img_u8, size, size), mean, std)
