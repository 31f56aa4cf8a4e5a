"""Image <-> tensor helpers on either side of the hot path (reference: basicsr/utils/img_util.py:9-153,
used by inference_femasr.py:54-67).  Behaviour kept: BGR uint8/float HWC <-> RGB float CHW, clamp to
min_max, x255 + round for uint8, BGR on the way out."""
from __future__ import annotations

import os

import numpy as np
import torch


def _one_to_tensor(img: np.ndarray, bgr2rgb: bool, float32: bool) -> torch.Tensor:
    if img.ndim == 2:
        img = img[:, :, None]
    if img.shape[2] == 3 and bgr2rgb:
        if img.dtype == np.float64:
            img = img.astype(np.float32)
        img = img[:, :, ::-1]
    t = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
    return t.float() if float32 else t


def img2tensor(imgs, bgr2rgb=True, float32=True):
    """HWC ndarray (BGR) or list of them -> CHW tensor(s) (RGB)."""
    if isinstance(imgs, list):
        return [_one_to_tensor(i, bgr2rgb, float32) for i in imgs]
    return _one_to_tensor(imgs, bgr2rgb, float32)


def _grid(t: torch.Tensor) -> torch.Tensor:
    """Tile a [N,C,H,W] batch into one image, sqrt(N) per row, 2px padding (torchvision make_grid layout)."""
    n, c, h, w = t.shape
    if n == 1:
        return t[0]
    if c == 1:
        t = t.expand(n, 3, h, w)
        c = 3
    ncol = min(int(np.sqrt(n)), n)
    nrow = int(np.ceil(n / ncol))
    pad = 2
    canvas = t.new_zeros((c, nrow * (h + pad) + pad, ncol * (w + pad) + pad))
    for k in range(n):
        r, q = divmod(k, ncol)
        y, x = r * (h + pad) + pad, q * (w + pad) + pad
        canvas[:, y:y + h, x:x + w] = t[k]
    return canvas


def tensor2img(tensor, rgb2bgr=True, out_type=np.uint8, min_max=(0, 1)):
    """[B,3|1,H,W] / [3|1,H,W] / [H,W] RGB tensor(s) -> HWC (BGR) / HW ndarray(s)."""
    single = torch.is_tensor(tensor)
    if not (single or (isinstance(tensor, list) and all(torch.is_tensor(t) for t in tensor))):
        raise TypeError(f"tensor or list of tensors expected, got {type(tensor)}")
    lo, hi = min_max
    outs = []
    for t in ([tensor] if single else tensor):
        t = t.squeeze(0).float().detach().cpu().clamp_(lo, hi)
        t = (t - lo) / (hi - lo)
        if t.dim() == 4:
            t = _grid(t)
        if t.dim() == 3:
            arr = t.numpy().transpose(1, 2, 0)
            if arr.shape[2] == 1:
                arr = arr[:, :, 0]
            elif rgb2bgr:
                arr = arr[:, :, ::-1]
        elif t.dim() == 2:
            arr = t.numpy()
        else:
            raise TypeError(f"Only support 4D, 3D or 2D tensor. But received with dimension: {t.dim()}")
        if out_type == np.uint8:
            arr = (arr * 255.0).round()
        outs.append(np.ascontiguousarray(arr.astype(out_type)))
    return outs[0] if len(outs) == 1 else outs


def imwrite(img, file_path, params=None, auto_mkdir=True):
    import cv2
    if auto_mkdir:
        os.makedirs(os.path.abspath(os.path.dirname(file_path)), exist_ok=True)
    if not cv2.imwrite(file_path, img, params):
        raise IOError("Failed in writing images.")


def imfrombytes(content, flag="color", float32=False):
    import cv2
    flags = {"color": cv2.IMREAD_COLOR, "grayscale": cv2.IMREAD_GRAYSCALE, "unchanged": cv2.IMREAD_UNCHANGED}
    img = cv2.imdecode(np.frombuffer(content, np.uint8), flags[flag])
    return img.astype(np.float32) / 255.0 if float32 else img


def crop_border(imgs, crop_border):
    if crop_border == 0:
        return imgs
    cut = (lambda v: v[crop_border:-crop_border, crop_border:-crop_border, ...])
    return [cut(v) for v in imgs] if isinstance(imgs, list) else cut(imgs)
