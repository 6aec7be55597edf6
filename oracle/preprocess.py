"""Input-pipeline oracle (TEST INFRASTRUCTURE ONLY): the reference's ``read_images`` / ``DiscoGANDataset`` image
arithmetic (``dataset.py:37-73,194-261``) twice --

* ``read_images_cv2``: the reference's own statement order on its own third-party dependencies (PIL decode, numpy,
  ``cv2.dilate`` / ``cv2.resize``; opencv is un-pinned in ``requirements.txt``, 4.13 is what this image has);
* ``preprocess_restated``: the published algorithm behind those cv2 calls restated in numpy integer / float64
  arithmetic (OpenCV ``resize.cpp``: 11-bit fixed-point bilinear for uint8, float coefficients for float64; 3x3
  ``dilate`` with the default border = maximum over the in-bounds neighbourhood).

tests/test_oracle.py pins the restatement to cv2 bit for bit on seeded images; tests/test_dataset_gpu.py holds the CUDA
kernel (``dg_preprocess_u8``) to both, bit for bit.
"""
import numpy as np


def domain_crop(domain, width):
    """(x0, crop_width, mode) of dataset.py:52-60: 'A' = left 256 columns + edge thickening, 'B' = columns from 256."""
    if domain == "A":
        return 0, min(256, width), 1
    if domain == "B":
        return 256, width - 256, 0
    return 0, width, 0


def _coeffs(src, dst, clamp):
    scale = 1.0 / (np.float64(dst) / np.float64(src))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0, src - 1
    return s, f


def _erode3(img):
    """255 - dilate3x3(255 - img): minimum over the in-bounds 3x3 neighbourhood."""
    H, W, _ = img.shape
    big = np.pad(img.astype(np.int64), ((1, 1), (1, 1), (0, 0)), constant_values=1 << 20)
    out = np.full(img.shape, 1 << 20, dtype=np.int64)
    for dy in range(3):
        for dx in range(3):
            out = np.minimum(out, big[dy:dy + H, dx:dx + W])
    return out


def preprocess_restated(image_u8, domain, image_size):
    """uint8 [H,W,3] -> float32 [3,S,S], the arithmetic of dataset.py:50-66 without cv2."""
    x0, cw, mode = domain_crop(domain, image_u8.shape[1])
    img = image_u8[:, x0:x0 + cw, :]
    H, W, _ = img.shape
    S = image_size
    sx, fx = _coeffs(W, S, True)
    sy, fy = _coeffs(H, S, False)
    x1 = np.minimum(sx + 1, W - 1)
    r0, r1 = np.clip(sy, 0, H - 1), np.clip(sy + 1, 0, H - 1)
    one = np.float32(1)
    if mode == 0:
        src = img.astype(np.int64)
        a0 = np.rint((one - fx) * np.float32(2048)).astype(np.int64)[None, :, None]
        a1 = np.rint(fx * np.float32(2048)).astype(np.int64)[None, :, None]
        b0 = np.rint((one - fy) * np.float32(2048)).astype(np.int64)[:, None, None]
        b1 = np.rint(fy * np.float32(2048)).astype(np.int64)[:, None, None]
        h0 = src[r0][:, sx] * a0 + src[r0][:, x1] * a1
        h1 = src[r1][:, sx] * a0 + src[r1][:, x1] * a1
        out = ((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2).clip(0, 255).astype(np.uint8)
        res = out.astype(np.float32) / 255.
    else:
        src = _erode3(img).astype(np.float64)
        a0, a1 = (one - fx).astype(np.float64)[None, :, None], fx.astype(np.float64)[None, :, None]
        b0, b1 = (one - fy).astype(np.float64)[:, None, None], fy.astype(np.float64)[:, None, None]
        h0 = src[r0][:, sx] * a0 + src[r0][:, x1] * a1
        h1 = src[r1][:, sx] * a0 + src[r1][:, x1] * a1
        res = (h0 * b0 + h1 * b1).astype(np.float32) / 255.
    return np.ascontiguousarray(res.transpose(2, 0, 1))


def preprocess_cv2(image_u8, domain, image_size):
    """dataset.py:50-66 verbatim in statement order, on an already decoded uint8 array."""
    import cv2
    image = image_u8
    if domain == "A":
        kernel = np.ones((3, 3), np.uint8)
        image = image[:, :256, :]
        image = 255. - image
        image = cv2.dilate(image, kernel, iterations=1)
        image = 255. - image
    elif domain == "B":
        image = image[:, 256:, :]
    image = cv2.resize(image, (image_size, image_size))
    image = image.astype(np.float32) / 255.
    return image.transpose(2, 0, 1)


def read_images_cv2(filenames, domain=None, image_size=64):
    """dataset.py:37-73: decode with PIL, preprocess, stack."""
    from PIL import Image
    images = [preprocess_cv2(np.array(Image.open(fn).convert("RGB")), domain, image_size) for fn in filenames]
    if not images:
        raise ValueError("no valid images")
    return np.stack(images)
