"""Helpers for the -m gpu parity tests: thin torch wrappers over the C-ABI stage entry points."""
import ctypes

import numpy as np
import torch

from comfystereo_b200 import _lib, engine
from comfystereo_b200._lib import CsParams, FILL_KEYS


def dev():
    return torch.device("cuda", 0)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rgbx(img_u8):
    """[...,3] uint8 -> [...,4] uint8 (X = 0), the library's packed pixel format."""
    out = np.zeros(img_u8.shape[:-1] + (4,), np.uint8)
    out[..., :3] = img_u8
    return out


def blur(d255, strength, thr, falloff, vert):
    """cs_blur on [n,h,w] or [h,w] float32 (0..255 scale) -> (L, R, minmax[n,4]) numpy."""
    d = np.ascontiguousarray(d255, np.float32)
    single = d.ndim == 2
    if single:
        d = d[None]
    n, h, w = d.shape
    p = CsParams()
    p.blur_enabled = 1
    p.blur_box = int(round(float(strength)))
    p.blur_radius = int(strength)
    p.blur_vert_smooth = int(vert)
    p.blur_edge_threshold = float(thr)
    p.blur_falloff = float(falloff)
    t = torch.from_numpy(d).to(dev())
    L, R = torch.empty_like(t), torch.empty_like(t)
    mm = torch.empty((n, 4), dtype=torch.float32, device=dev())
    scratch = torch.empty(2 * n * h * w, dtype=torch.uint8, device=dev())
    _lib.check(_lib.lib().cs_blur(t.data_ptr(), n, h, w, ctypes.byref(p), L.data_ptr(), R.data_ptr(),
                                  mm.data_ptr(), scratch.data_ptr(), stream()))
    torch.cuda.synchronize()
    L, R = L.cpu().numpy(), R.cpu().numpy()
    return (L[0], R[0], mm.cpu().numpy()) if single else (L, R, mm.cpu().numpy())


def depth_resize(depth, size):
    """cs_depth_resize: depth [n,dh,dw,c] float32 -> gray [n,h,w] at size=(h,w)."""
    d = np.ascontiguousarray(depth, np.float32)
    n, dh, dw, c = d.shape
    h, w = int(size[0]), int(size[1])
    t = torch.from_numpy(d).to(dev())
    out = torch.empty((n, h, w), dtype=torch.float32, device=dev())
    _lib.check(_lib.lib().cs_depth_resize(t.data_ptr(), n, dh, dw, c, h, w, out.data_ptr(), stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def warp_fill(img_u8, depth, fill_key, divergence, separation, expo, conv, exact=False, flags=0):
    """cs_warp_fill on ONE eye: img_u8 [n,h,w,3] or [h,w,3], depth same leading dims.  Returns uint8 [...,4]
    (RGB + the filled/mask flag byte)."""
    img = np.ascontiguousarray(img_u8, np.uint8)
    d = np.ascontiguousarray(depth, np.float32)
    single = d.ndim == 2
    if single:
        img, d = img[None], d[None]
    n, h, w = d.shape
    ti = torch.from_numpy(rgbx(img)).to(dev())
    td = torch.from_numpy(d).to(dev())
    out = torch.empty_like(ti)
    lib = _lib.lib()
    nb = lib.cs_warp_fill_scratch_bytes(n, h, w)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev())
    lib.cs_set_test_flags((1 if exact else 0) | flags)
    try:
        _lib.check(lib.cs_warp_fill(ti.data_ptr(), td.data_ptr(), n, h, w, FILL_KEYS.index(fill_key),
                                    float(divergence), float(separation), float(expo), float(conv),
                                    out.data_ptr(), scratch.data_ptr(), nb, stream()))
        torch.cuda.synchronize()
    finally:
        lib.cs_set_test_flags(0)
    o = out.cpu().numpy()
    return o[0] if single else o


def shift_indices(nd, div_px, sep_px, expo, kind):
    t = torch.from_numpy(np.ascontiguousarray(nd, np.float32)).to(dev())
    h, w = t.shape
    out = torch.empty((h, w), dtype=torch.int32, device=dev())
    _lib.check(_lib.lib().cs_shift_indices(t.data_ptr(), 1, h, w, float(div_px), float(sep_px), float(expo),
                                           int(kind), out.data_ptr(), stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def forward_warp(img_hwc, depth, div_px, sep_px, expo, conv):
    """cs_forward_warp: img [n,h,w,3] float32, depth [n,h,w] -> (warped [n,h,w,3], mask bool [n,h,w])."""
    ti = torch.from_numpy(np.ascontiguousarray(img_hwc, np.float32)).to(dev())
    td = torch.from_numpy(np.ascontiguousarray(depth, np.float32)).to(dev())
    n, h, w = td.shape
    out = torch.empty_like(ti)
    mask = torch.empty((n, h, w), dtype=torch.float32, device=dev())
    scratch = torch.empty(_lib.lib().cs_forward_warp_scratch_bytes(n, h, w), dtype=torch.uint8, device=dev())
    _lib.check(_lib.lib().cs_forward_warp(ti.data_ptr(), td.data_ptr(), n, h, w, float(div_px), float(sep_px),
                                          float(expo), float(conv), out.data_ptr(), mask.data_ptr(),
                                          scratch.data_ptr(), scratch.numel(), stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy(), mask.cpu().numpy() > 0.5


def compose(left_u8, right_u8, mode):
    tl = torch.from_numpy(rgbx(np.ascontiguousarray(left_u8))[None]).to(dev())
    tr = torch.from_numpy(rgbx(np.ascontiguousarray(right_u8))[None]).to(dev())
    _, h, w, _ = tl.shape
    m = engine.MODES.index(mode)
    ho, wo = (h, 2 * w) if m in (0, 1) else ((2 * h, w) if m in (2, 3) else (h, w))
    st = torch.empty((1, ho, wo, 3), dtype=torch.float32, device=dev())
    mk = torch.empty((1, ho, wo), dtype=torch.float32, device=dev())
    _lib.check(_lib.lib().cs_compose(tl.data_ptr(), tr.data_ptr(), 1, h, w, m, st.data_ptr(), mk.data_ptr(), stream()))
    torch.cuda.synchronize()
    return st[0].cpu().numpy(), mk[0].cpu().numpy()


def q8(a):
    return np.rint(np.asarray(a) * 255.0).astype(np.uint8)
