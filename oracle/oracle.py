"""ctypes front-end of the CPU oracle (oracle/stereo_oracle.c) + numpy restatement of the glue.

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this.

Glue restated here (numpy, per frame) with the reference lines it follows:
  * StereoImageNode.generate prep / result conversion        GS:117-269, GS:355-378
  * create_stereoimages (CPU techniques)                      SIG:1422-1574
  * create_stereoimages_gpu (GPU Warp, scatter variant)       SIG:1005-1128
The per-pixel algorithms live in the C file.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstereo_oracle.so")

MODES = ["left-right", "right-left", "top-bottom", "bottom-top", "red-cyan-anaglyph",
         "left-only", "only-right", "cyan-red-reverseanaglyph"]

FILL_NAME_TO_KEY = {  # GS:88-100
    'GPU Warp (Fast)': 'gpu_warp',
    'No fill': 'none',
    'No fill - Reverse projection': 'inverse',
    'Imperfect fill - Hybrid Edge': 'hybrid_edge',
    'Fill - Naive': 'naive',
    'Fill - Naive interpolating': 'naive_interpolating',
    'Fill - Polylines Soft': 'polylines_soft',
    'Fill - Polylines Sharp': 'polylines_sharp',
    'Fill - Post-fill': 'none_post',                              # GS:97-99: mapped but commented out of the dropdown
    'Fill - Reverse projection with Post-fill': 'inverse_post',
    'Fill - Hybrid Edge with fill': 'hybrid_edge_plus',
}


def build(force=False):
    if force or not os.path.isfile(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "stereo_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_abi_version.restype = ctypes.c_int
    return _lib


def set_threads(n=0):
    """Sets (n > 0) and returns the number of host threads the oracle's OpenMP loops use."""
    f = lib().orc_set_threads
    f.restype = ctypes.c_int
    return int(f(ctypes.c_int(int(n))))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------ thin wrappers
def gray3(rgb):
    rgb = _c(rgb, np.float32)
    out = np.empty(rgb.shape[:-1], np.float32)
    lib().orc_gray3(_p(rgb), ctypes.c_int64(out.size), _p(out))
    return out


def resize_bilinear(gray, size):
    """N1 (GS:141-148 / GS:214-220): F.interpolate(gray[None, None], size, 'bilinear', align_corners=False) on a
    [H,W] or [N,H,W] float32 array, strict float32 arithmetic (see stereo_oracle.c)."""
    g = _c(gray, np.float32)
    if g.ndim == 3:
        return np.stack([resize_bilinear(f, size) for f in g])
    oh, ow = int(size[0]), int(size[1])
    out = np.empty((oh, ow), np.float32)
    lib().orc_resize_bilinear(_p(g), g.shape[0], g.shape[1], oh, ow, _p(out))
    return out


def py_round_half_even(x):
    return int(round(float(x)))  # python round() == banker's rounding, SIG:1208 (Q11)


def blur(depth255, strength, edge_threshold, falloff=1.0, vert_smooth=0):
    """directional_motion_blur_gpu(depth, s, thr, s, falloff, vert)  SIG:1171-1251."""
    d = _c(depth255, np.float32)
    if strength <= 0:
        return d, d
    bs, R = py_round_half_even(strength), int(strength)
    if bs <= 0:
        raise RuntimeError("kernel size should be greater than zero")  # conv2d's error, Q11
    H, W = d.shape
    L, Rr = np.empty_like(d), np.empty_like(d)
    rc = lib().orc_blur(_p(d), H, W, bs, R, ctypes.c_float(edge_threshold), ctypes.c_float(falloff),
                        int(vert_smooth), _p(L), _p(Rr))
    assert rc == 0
    return L, Rr


def normalize(depth, conv):
    d = _c(depth, np.float32)
    nd = np.empty_like(d)
    lib().orc_normalize(_p(d), ctypes.c_int64(d.size), ctypes.c_float(np.float32(conv)), _p(nd))
    return nd


def shift_indices(nd, div_px, sep_px, expo, kind):
    nd = _c(nd, np.float32)
    H, W = nd.shape
    out = np.empty((H, W), np.int32)
    lib().orc_shift_indices(_p(nd), H, W, ctypes.c_double(div_px), ctypes.c_double(sep_px),
                            ctypes.c_double(expo), int(kind), _p(out))
    return out


def _warp_args(img, nd, div_px, sep_px, expo):
    img = _c(img, np.uint8)
    nd = _c(nd, np.float32)
    H, W = nd.shape
    assert img.shape == (H, W, 3)
    return img, nd, H, W, ctypes.c_double(div_px), ctypes.c_double(sep_px), ctypes.c_double(expo)


def naive(img, nd, div_px, sep_px, expo, fill, want_src=False):
    img, nd, H, W, a, b, c = _warp_args(img, nd, div_px, sep_px, expo)
    out = np.empty_like(img)
    src = np.empty((H, W), np.int32) if want_src else None
    code = {'none': 0, 'naive': 1, 'naive_interpolating': 2}[fill]
    lib().orc_naive(_p(img), _p(nd), H, W, a, b, c, code, _p(out), _p(src) if want_src else None)
    return (out, src) if want_src else out


def polylines(img, nd, div_px, sep_px, expo, sharp):
    img, nd, H, W, a, b, c = _warp_args(img, nd, div_px, sep_px, expo)
    out = np.empty_like(img)
    lib().orc_polylines(_p(img), _p(nd), H, W, a, b, c, int(bool(sharp)), _p(out))
    return out


def inverse(img, nd, div_px, sep_px, expo, want_mask=False):
    img, nd, H, W, a, b, c = _warp_args(img, nd, div_px, sep_px, expo)
    out = np.empty_like(img)
    mask = np.empty((H, W), np.uint8)
    lib().orc_inverse(_p(img), _p(nd), H, W, a, b, c, _p(out), _p(mask))
    return (out, mask) if want_mask else out


def hybrid_stage1(img, nd, div_px, sep_px, expo):
    img, nd, H, W, a, b, c = _warp_args(img, nd, div_px, sep_px, expo)
    out = np.empty_like(img)
    mask = np.empty((H, W), np.uint8)
    lib().orc_hybrid_stage1(_p(img), _p(nd), H, W, a, b, c, _p(out), _p(mask))
    return out, mask


def hybrid_edge(img, nd, div_px, sep_px, expo):
    base, mask = hybrid_stage1(img, nd, div_px, sep_px, expo)
    img = _c(img, np.uint8)
    H, W = mask.shape
    out = np.empty_like(base)
    lib().orc_hybrid_stage2(_p(img), _p(base), _p(mask), H, W, _p(out))
    return out


def interp_rows(base, mask):
    """np.interp row fill of the *_post variants, SIG:1804-1833."""
    base, mask = _c(base, np.uint8), _c(mask, np.uint8)
    H, W = mask.shape
    out = np.empty_like(base)
    lib().orc_interp_rows(_p(base), _p(mask), H, W, _p(out))
    return out


def naive_post(img, nd, div_px, sep_px, expo):
    base, src = naive(img, nd, div_px, sep_px, expo, 'none', want_src=True)
    return interp_rows(base, (src >= 0).astype(np.uint8))


def inverse_post(img, nd, div_px, sep_px, expo):
    base, mask = inverse(img, nd, div_px, sep_px, expo, want_mask=True)
    return interp_rows(base, mask)


def hybrid_edge_plus(img, nd, div_px, sep_px, expo):
    prim = hybrid_edge(img, nd, div_px, sep_px, expo)
    poly = polylines(img, nd, div_px, sep_px, expo, False)
    out = np.empty_like(prim)
    lib().orc_merge_black(_p(prim), _p(poly), ctypes.c_int64(prim.shape[0] * prim.shape[1]), _p(out))
    return out


def compose_u8(left, right, mode):
    left, right = _c(left, np.uint8), _c(right, np.uint8)
    H, W, _ = left.shape
    m = MODES.index(mode)
    shape = (H, 2 * W, 3) if m in (0, 1) else ((2 * H, W, 3) if m in (2, 3) else (H, W, 3))
    out = np.empty(shape, np.uint8)
    lib().orc_compose_u8(_p(left), _p(right), H, W, m, _p(out))
    return out


def gpuwarp_eye(img_chw, depth01, div_px, sep_px, expo, conv):
    """forward_warp_gpu for ONE frame; depth01 already /255 (SIG:313-316 handled by caller)."""
    img = _c(img_chw, np.float32)
    d = _c(depth01, np.float32)
    H, W = d.shape
    out = np.empty_like(img)
    mask = np.empty((H, W), np.uint8)
    f = lambda v: ctypes.c_float(np.float32(v))
    lib().orc_gpuwarp(_p(img), _p(d), H, W, f(div_px), f(sep_px), f(expo), f(conv), _p(out), _p(mask))
    return out, mask.astype(bool)


def meshwarp_batch(img_bchw, depth01_bhw, div_px, sep_px, expo, conv):
    """forward_warp_mesh (SIG:453-689) for a whole sub-batch; depth01 already /255 (SIG:487-489 handled by the caller).
    The triangle culling is batch-wide (SIG:522-537: kept when it passes in ANY frame), the rasterisation is
    orc_mesh_raster's rule set -- OpenGL's own is implementation-defined, so parity with the reference is UNPINNED here."""
    img = _c(img_bchw, np.float32)
    d = _c(depth01_bhw, np.float32)
    B, H, W = d.shape
    f = lambda v: ctypes.c_float(np.float32(v))
    nd = np.empty((B, H, W), np.float32)
    po = np.empty((B, H, W), np.float32)
    for b in range(B):
        lib().orc_mesh_offsets(_p(d[b]), H, W, f(div_px), f(sep_px), f(expo), f(conv), _p(nd[b]), _p(po[b]))
    out = np.zeros_like(img)
    mask = np.ones((B, H, W), np.uint8)
    if H < 2 or W < 2:      # no triangles: nothing is covered and there is nothing to smear from
        return out, mask.astype(bool)
    o00, o10, o01, o11 = po[:, :-1, :-1], po[:, :-1, 1:], po[:, 1:, :-1], po[:, 1:, 1:]
    thr = np.float32(1.5)
    with np.errstate(invalid='ignore'):
        diag = np.abs(o10 - o01)
        ka = np.maximum(np.maximum(np.abs(o00 - o10), np.abs(o00 - o01)), diag) < thr
        kb = np.maximum(np.maximum(np.abs(o11 - o10), np.abs(o11 - o01)), diag) < thr
    keep_a = _c(ka.any(axis=0), np.uint8)
    keep_b = _c(kb.any(axis=0), np.uint8)
    for b in range(B):
        lib().orc_mesh_raster(_p(img[b]), _p(nd[b]), _p(po[b]), _p(keep_a), _p(keep_b), H, W,
                              1 if np.float32(div_px) >= 0 else 0, _p(out[b]), _p(mask[b]))
    return out, mask.astype(bool)


# ------------------------------------------------------------------ pipeline glue
def apply_stereo_divergence(img_u8, depth, divergence, separation, expo, fill, conv):
    """SIG:1576-1620."""
    assert img_u8.shape[:2] == depth.shape, 'Depthmap and the image must have the same size'
    nd = normalize(depth, conv)
    W = img_u8.shape[1]
    div_px = (divergence / 100.0) * W
    sep_px = (separation / 100.0) * W
    if fill in ('none', 'naive', 'naive_interpolating'):
        return naive(img_u8, nd, div_px, sep_px, expo, fill)
    if fill in ('polylines_soft', 'polylines_sharp'):
        return polylines(img_u8, nd, div_px, sep_px, expo, fill == 'polylines_sharp')
    if fill == 'inverse':
        return inverse(img_u8, nd, div_px, sep_px, expo)
    if fill == 'hybrid_edge':
        return hybrid_edge(img_u8, nd, div_px, sep_px, expo)
    if fill == 'none_post':
        return naive_post(img_u8, nd, div_px, sep_px, expo)
    if fill == 'inverse_post':
        return inverse_post(img_u8, nd, div_px, sep_px, expo)
    if fill == 'hybrid_edge_plus':
        return hybrid_edge_plus(img_u8, nd, div_px, sep_px, expo)
    return img_u8  # SIG:1620 fallback


def wrap_depth_u8(depth255):
    """(x * 255).astype(uint8) on an x that is already 0..255 (Q1): wraps mod 256, SIG:1511-1516."""
    v = (np.asarray(depth255, np.float32) * np.float32(255)).astype(np.int64)
    return (v % 256).astype(np.uint8)


def create_stereoimages(image_chw, depth_hw, divergence, separation=0.0, modes=None,
                        stereo_balance=0.0, stereo_offset_exponent=1.0, fill_technique='polylines_sharp',
                        depth_blur_strength=0.0, depth_blur_edge_threshold=6.0,
                        direction_aware_depth_blur=False, convergence_point=0.5,
                        depth_blur_falloff=1.0, depth_blur_vert_smooth=0, blur_override=None):
    """Tensor-input branch of create_stereoimages (SIG:1466-1574) on numpy arrays.
    Returns (list of composed uint8 images, left depth uint8, right depth uint8)."""
    if modes is None:
        modes = ['left-right']
    if not isinstance(modes, list):
        modes = [modes]
    depth = np.ascontiguousarray(depth_hw, np.float32)
    if depth.max() <= 1.0:
        depth = depth * np.float32(255.0)
    if direction_aware_depth_blur:
        if blur_override is not None:  # stage-wise parity: the reference's own blurred depth
            dl, dr = (np.ascontiguousarray(b, np.float32) for b in blur_override)
        else:
            dl, dr = blur(depth, depth_blur_strength, depth_blur_edge_threshold,
                          depth_blur_falloff, depth_blur_vert_smooth)
    else:
        dl = dr = depth
    img = np.asarray(image_chw, np.float32)
    if img.ndim == 3 and img.shape[0] == 3:
        img = img.transpose(1, 2, 0)
    img_u8 = np.clip(img * np.float32(255), 0, 255).astype(np.uint8)  # truncates (Q2)
    mod_l, mod_r = wrap_depth_u8(dl), wrap_depth_u8(dr)
    ldiv = divergence * (1 + stereo_balance)
    rdiv = divergence * (1 - stereo_balance)
    left = img_u8 if ldiv < 0.001 else apply_stereo_divergence(
        img_u8, dl, +1 * ldiv, -1 * separation, stereo_offset_exponent, fill_technique, convergence_point)
    right = img_u8 if rdiv < 0.001 else apply_stereo_divergence(
        img_u8, dr, -1 * rdiv, separation, stereo_offset_exponent, fill_technique, convergence_point)
    results = []
    for mode in modes:
        if mode not in MODES:
            raise Exception('Unknown mode')
        results.append(compose_u8(left, right, mode))
    return results, mod_l, mod_r


def create_stereoimages_gpu(image_bchw, depth_bhw, divergence, separation=0.0, modes=None,
                            stereo_balance=0.0, stereo_offset_exponent=1.0, convergence_point=0.5,
                            depth_blur_strength=0.0, depth_blur_edge_threshold=6.0,
                            direction_aware_depth_blur=False, depth_blur_falloff=1.0,
                            depth_blur_vert_smooth=0, blur_override=None, mesh=False):
    """SIG:1005-1128, one sub-batch; warp_fn = forward_warp_gpu (moderngl absent) or, with mesh=True, forward_warp_mesh."""
    if modes is None:
        modes = ['left-right']
    if not isinstance(modes, list):
        modes = [modes]
    img = np.ascontiguousarray(image_bchw, np.float32)
    depth = np.ascontiguousarray(depth_bhw, np.float32)
    B, _, H, W = img.shape
    if depth.max() <= 1.0:  # sub-batch-wide (Q9)
        depth = depth * np.float32(255.0)
    if direction_aware_depth_blur and depth_blur_strength > 0 and blur_override is not None:
        dl, dr = (np.ascontiguousarray(b, np.float32) for b in blur_override)
    elif direction_aware_depth_blur and depth_blur_strength > 0:
        pairs = [blur(depth[b], depth_blur_strength, depth_blur_edge_threshold,
                      depth_blur_falloff, depth_blur_vert_smooth) for b in range(B)]
        dl = np.stack([p[0] for p in pairs])
        dr = np.stack([p[1] for p in pairs])
    else:
        dl = dr = depth
    ldiv = divergence * (1 + stereo_balance)
    rdiv = divergence * (1 - stereo_balance)
    ldiv_px, rdiv_px, sep_px = (ldiv / 100.0) * W, (rdiv / 100.0) * W, (separation / 100.0) * W

    def warp(depth_b, div_px, sp):
        d = depth_b
        if (d.reshape(B, -1).max(axis=1) > 1.0).any():  # SIG:314-316, whole sub-batch
            d = d / np.float32(255.0)
        if mesh:
            return meshwarp_batch(img, d, div_px, sp, stereo_offset_exponent, convergence_point)
        outs, masks = [], []
        for b in range(B):
            o, m = gpuwarp_eye(img[b], d[b], div_px, sp, stereo_offset_exponent, convergence_point)
            outs.append(o)
            masks.append(m)
        return np.stack(outs), np.stack(masks)

    lmask = np.zeros((B, H, W), bool)
    rmask = np.zeros((B, H, W), bool)
    left, right = img, img
    if not ldiv < 0.001:
        left, lmask = warp(dl, +ldiv_px, -sep_px)
    if not rdiv < 0.001:
        right, rmask = warp(dr, -rdiv_px, sep_px)
    results = []
    for mode in modes:
        if mode == 'left-right':
            r = np.concatenate([left, right], axis=3)
        elif mode == 'right-left':
            r = np.concatenate([right, left], axis=3)
        elif mode == 'top-bottom':
            r = np.concatenate([left, right], axis=2)
        elif mode == 'bottom-top':
            r = np.concatenate([right, left], axis=2)
        elif mode == 'red-cyan-anaglyph':
            r = np.stack([left[:, 0], right[:, 1], right[:, 2]], axis=1)
        elif mode == 'left-only':
            r = left
        elif mode == 'only-right':
            r = right
        elif mode == 'cyan-red-reverseanaglyph':
            r = np.stack([right[:, 0], left[:, 1], left[:, 2]], axis=1)
        else:
            raise ValueError(f'Unknown mode: {mode}')
        results.append(r)
    dlo = dl / np.float32(255.0) if dl.max() > 1.0 else dl
    dro = dr / np.float32(255.0) if dr.max() > 1.0 else dr
    return results, dlo, dro, (lmask | rmask)


def node_generate(image, depth_map, divergence=4.5, separation=0.0, modes="left-right",
                  stereo_balance=0.0, convergence_point=0.5, stereo_offset_exponent=2.0,
                  fill_technique='GPU Warp (Fast)', depth_blur_edge_threshold=20.0,
                  depth_blur_strength=20.0, depth_map_blur=True, depth_blur_falloff=1.0,
                  depth_blur_vert_smooth=0, batch_size=4, blur_override=None, mesh=False):
    """StereoImageNode.generate (GS:79-353) on numpy arrays: image [N,H,W,3], depth [N,H,W,C].
    Returns (stereo [N,Ho,Wo,3], depth_left [N,H,W,3], depth_right [N,H,W,3], mask [N,Hm,Wm]) float32."""
    image = np.asarray(image, np.float32)
    depth_map = np.asarray(depth_map, np.float32)
    key = FILL_NAME_TO_KEY.get(fill_technique, 'gpu_warp')
    N = image.shape[0]
    f255 = np.float32(255.0)
    st, dls, drs, ms = [], [], [], []
    if key == 'gpu_warp':
        gb = min(batch_size, N)
        for s in range(0, N, gb):
            img = image[s:s + gb].transpose(0, 3, 1, 2)
            dm = depth_map[s:s + gb]
            if dm.shape[3] == 3:
                dm = gray3(dm)
            else:
                dm = dm[..., 0]
            if dm.shape[1:] != img.shape[2:]:                      # GS:141-148
                dm = resize_bilinear(dm, img.shape[2:])
            res, dl, dr, mask = create_stereoimages_gpu(
                img, dm, divergence, separation, [modes], stereo_balance, stereo_offset_exponent,
                convergence_point, depth_blur_strength, depth_blur_edge_threshold, depth_map_blur,
                depth_blur_falloff, depth_blur_vert_smooth,
                blur_override=None if blur_override is None else tuple(b[s:s + gb] for b in blur_override), mesh=mesh)
            st.append(res[0].transpose(0, 2, 3, 1))
            dls.append(np.repeat(np.clip(dl, 0, 1)[..., None], 3, axis=-1))
            drs.append(np.repeat(np.clip(dr, 0, 1)[..., None], 3, axis=-1))
            ms.append(mask.astype(np.float32))
    else:
        for i in range(N):
            dm = depth_map[i]
            dm = gray3(dm) if dm.shape[2] == 3 else dm[..., 0]
            if dm.shape != image[i].shape[:2]:                    # GS:214-220
                dm = resize_bilinear(dm, image[i].shape[:2])
            res, ml, mr = create_stereoimages(
                image[i].transpose(2, 0, 1), dm, divergence, separation, [modes], stereo_balance,
                stereo_offset_exponent, key, depth_blur_strength, depth_blur_edge_threshold,
                depth_map_blur, convergence_point, depth_blur_falloff, depth_blur_vert_smooth,
                blur_override=None if blur_override is None else tuple(b[i] for b in blur_override))
            r = res[0]
            st.append((r.astype(np.float32) / f255)[None])
            dls.append(np.repeat((ml.astype(np.float32) / f255)[None, ..., None], 3, axis=-1))
            drs.append(np.repeat((mr.astype(np.float32) / f255)[None, ..., None], 3, axis=-1))
            black = (r.astype(np.int32).sum(axis=-1) == 0).astype(np.uint8) * 255  # GS:355-361 (Q6)
            ms.append((black.astype(np.float32) / f255)[None])
    return (np.concatenate(st), np.concatenate(dls), np.concatenate(drs), np.concatenate(ms))


# ------------------------------------------------------------------ non-tensor inputs (numpy / PIL): SIG:1346-1419
def _corr1d_f32(a, weights, axis, mode, origin=0):
    """scipy.ndimage.correlate1d on a float32 array with float64 weights, restated: every output sample is a float64
    sum over the (border-extended) line, rounded once to float32.  Odd symmetric / antisymmetric kernels take scipy's
    paired form (centre first, then pairs from the outside in); other kernels start with the LAST tap and then add the
    others in ascending order (NI_Correlate1D's general loop) -- the order matters: box sums of float32 samples hit exact
    float32 rounding ties about once in a thousand."""
    a = np.asarray(a, np.float32)
    w = np.asarray(weights, np.float64)
    n = len(w)
    s1 = n // 2 + origin
    s2 = n - n // 2 - 1 - origin
    pad = [(0, 0)] * a.ndim
    pad[axis] = (s1, s2)
    ext = np.pad(a, pad, mode={'reflect': 'symmetric', 'nearest': 'edge'}[mode]).astype(np.float64)
    L = a.shape[axis]

    def tap(i):   # extended line shifted so that tap i lines up with the output sample
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(i, i + L)
        return ext[tuple(sl)]

    half = n // 2
    sym = (n & 1) and all(abs(w[half - k] - w[half + k]) <= np.finfo(np.float64).eps for k in range(1, half + 1))
    asym = (n & 1) and all(abs(w[half - k] + w[half + k]) <= np.finfo(np.float64).eps for k in range(1, half + 1))
    if sym or asym:
        acc = tap(half) * w[half]
        for k in range(half, 0, -1):          # ll = -size1 .. -1: outermost pair first
            pair = (tap(half - k) + tap(half + k)) if sym else (tap(half + k) - tap(half - k))
            acc = acc + pair * (w[half - k] if sym else w[half + k])
    else:
        acc = tap(n - 1) * w[n - 1]
        for i in range(n - 1):
            acc = acc + tap(i) * w[i]
    return acc.astype(np.float32)


def blur_numpy(depth, strength, edge_threshold, falloff=1.0, vert_smooth=0):
    """directional_motion_blur(depth, s, thr, s, falloff, vert) -- the scipy blur the reference applies to NON-tensor
    inputs (SIG:1346-1419): Sobel with reflected borders, nearest-border box filters (scipy's convolve1d: an even box
    reaches one sample further right than left), float64 sums rounded to float32.  depth is used as given (no x255)."""
    d = np.ascontiguousarray(depth, np.float32)
    if strength <= 0:
        return d, d
    bs, R = py_round_half_even(strength), int(strength)
    h, w = d.shape
    g = _corr1d_f32(_corr1d_f32(d, [-1.0, 0.0, 1.0], 1, 'reflect'), [1.0, 2.0, 1.0], 0, 'reflect')   # scipy.ndimage.sobel
    e = np.clip(np.abs(g) / np.float32(10 * edge_threshold), 0, 1)
    masks = ((g > 0) & (e > 0.5), (g < 0) & (e > 0.5))
    cols = np.arange(w, dtype=np.float32)
    far = np.float32(R + 1)

    def weight(mask):
        last_l = np.maximum.accumulate(np.where(mask, cols[None, :], np.float32(-1)), axis=1)
        dist_l = np.where(last_l >= 0, cols[None, :] - last_l, far)
        last_r = np.maximum.accumulate(np.where(mask[:, ::-1], cols[None, :], np.float32(-1)), axis=1)
        dist_r = np.where(last_r >= 0, cols[None, :] - last_r, far)[:, ::-1]
        with np.errstate(divide='ignore', invalid='ignore'):
            wgt = np.clip(np.float32(1.0) - np.minimum(dist_l, dist_r) / np.float32(R), 0.0, 1.0) ** np.float32(falloff)
        return wgt.astype(np.float32)

    wl, wr = weight(masks[0]), weight(masks[1])
    if vert_smooth > 0:
        k = np.ones(2 * vert_smooth + 1) / (2 * vert_smooth + 1)
        wl = np.clip(_corr1d_f32(wl, k, 0, 'nearest'), 0.0, 1.0)
        wr = np.clip(_corr1d_f32(wr, k, 0, 'nearest'), 0.0, 1.0)
    box = np.ones(bs) / bs
    b = _corr1d_f32(d, box, 1, 'nearest', origin=(-1 if bs % 2 == 0 else 0))   # convolve1d: flipped box, origin - 1 if even
    one = np.float32(1.0)
    return (wl * b + (one - wl) * d).astype(np.float32), (wr * b + (one - wr) * d).astype(np.float32)


def create_stereoimages_arrays(image_u8, depth, divergence, separation=0.0, modes=None, stereo_balance=0.0,
                               stereo_offset_exponent=1.0, fill_technique='polylines_sharp', depth_blur_strength=0.0,
                               depth_blur_edge_threshold=6.0, direction_aware_depth_blur=False, convergence_point=0.5,
                               depth_blur_falloff=1.0, depth_blur_vert_smooth=0):
    """Non-tensor branch of create_stereoimages (SIG:1486-1496, 1520-1574): uint8 image and float depth used as given,
    scipy blur, depth outputs trunc(clip(depth, 0, 255)).  Returns (list of uint8 images, left depth u8, right depth u8)."""
    if modes is None:
        modes = ['left-right']
    if not isinstance(modes, list):
        modes = [modes]
    img = np.ascontiguousarray(image_u8, np.uint8)
    d = np.ascontiguousarray(depth, np.float32)
    if direction_aware_depth_blur:
        dl, dr = blur_numpy(d, depth_blur_strength, depth_blur_edge_threshold, depth_blur_falloff, depth_blur_vert_smooth)
    else:
        dl = dr = d
    ldiv, rdiv = divergence * (1 + stereo_balance), divergence * (1 - stereo_balance)
    left = img if ldiv < 0.001 else apply_stereo_divergence(
        img, dl, +1 * ldiv, -1 * separation, stereo_offset_exponent, fill_technique, convergence_point)
    right = img if rdiv < 0.001 else apply_stereo_divergence(
        img, dr, -1 * rdiv, separation, stereo_offset_exponent, fill_technique, convergence_point)
    results = []
    for mode in modes:
        if mode not in MODES:
            raise Exception('Unknown mode')
        results.append(compose_u8(left, right, mode))
    q = lambda a: np.clip(a, 0, 255).astype(np.uint8)
    return results, q(dl), q(dr)
