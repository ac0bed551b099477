"""Host side of the hot path: parameter marshalling, workspace management, device- and
host-buffer entry points, frame sharding across the GPUs of one box.

Everything numerical happens in libcomfystereo_b200.so; this module only allocates tensors with
torch and passes raw pointers and the current CUDA stream through the C ABI.
"""
import ctypes
import threading

import torch

from . import _lib
from ._lib import CsParams, FILL_KEYS, MODES

GPU_WARP_KEYS = ('gpu_warp', 'gpu_warp_mesh')

FILL_NAME_TO_KEY = {  # the node's dropdown labels -> dispatch keys, GS:88-100
    'GPU Warp (Fast)': 'gpu_warp',
    'No fill': 'none',
    'No fill - Reverse projection': 'inverse',
    'Imperfect fill - Hybrid Edge': 'hybrid_edge',
    'Fill - Naive': 'naive',
    'Fill - Naive interpolating': 'naive_interpolating',
    'Fill - Polylines Soft': 'polylines_soft',
    'Fill - Polylines Sharp': 'polylines_sharp',
    'Fill - Post-fill': 'none_post',                              # GS:97-99: still mapped, no longer in the dropdown
    'Fill - Reverse projection with Post-fill': 'inverse_post',
    'Fill - Hybrid Edge with fill': 'hybrid_edge_plus',
}


def make_params(fill_key, mode, divergence, separation=0.0, stereo_balance=0.0, convergence_point=0.5,
                stereo_offset_exponent=1.0, depth_blur=False, depth_blur_strength=0.0,
                depth_blur_edge_threshold=6.0, depth_blur_falloff=1.0, depth_blur_vert_smooth=0,
                group_size=0):
    """Widget values -> cs_params.  The two integers python derives from depth_blur_strength
    (SIG:1208-1209: bs = int(round(s)) with banker's rounding, R = int(s)) are computed here."""
    if mode not in MODES:
        raise ValueError(f'Unknown mode: {mode}')
    if fill_key not in FILL_KEYS:
        raise ValueError(f'Unknown fill technique key: {fill_key}')
    p = CsParams()
    p.fill = FILL_KEYS.index(fill_key)
    p.mode = MODES.index(mode)
    p.divergence = float(divergence)
    p.separation = float(separation)
    p.stereo_balance = float(stereo_balance)
    p.convergence_point = float(convergence_point)
    p.stereo_offset_exponent = float(stereo_offset_exponent)
    strength = float(depth_blur_strength)
    enabled = bool(depth_blur) and strength > 0  # strength <= 0 returns the input twice, SIG:1194
    p.blur_enabled = int(enabled)
    p.blur_box = int(round(strength)) if enabled else 0
    p.blur_radius = int(strength) if enabled else 0
    if enabled and p.blur_box <= 0:
        # what torch's conv2d says for the reference in this corner
        raise RuntimeError("kernel size should be greater than zero")
    p.blur_vert_smooth = int(depth_blur_vert_smooth)
    p.blur_edge_threshold = float(depth_blur_edge_threshold)
    p.blur_falloff = float(depth_blur_falloff)
    p.group_size = int(group_size)
    return p


def output_shapes(p, n, h, w):
    ho, wo, hm, wm = (ctypes.c_int() for _ in range(4))
    _lib.check(_lib.lib().cs_output_dims(ctypes.byref(p), h, w, ctypes.byref(ho), ctypes.byref(wo),
                                         ctypes.byref(hm), ctypes.byref(wm)))
    return (n, ho.value, wo.value, 3), (n, h, w, 3), (n, hm.value, wm.value)


def _workspace(device, nbytes):
    """Scratch for one call, from torch's caching allocator: it hands a block back only to allocations on the same
    stream, so concurrent calls (other threads, other streams) never share scratch and a block is never reused while
    kernels of an earlier call on another stream may still touch it."""
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def release():
    if _lib.loaded():
        _lib.lib().cs_host_release()


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def default_chunk(p, n, h, w):
    """Frames per kernel sequence: about sixteen 1080p frames' worth of pixels.  Measured on B200 (chunk sweep in
    profiles/): fewer, larger launches (fewer partial waves and launch gaps) matter more than keeping the ~26 B/px of
    scratch inside L2 -- 4 / 8 / 16 frames per sequence: 3.73 / 3.56 / 3.46 ms per 16 frames of 1080p Polylines Sharp."""
    group = p.group_size if (FILL_KEYS[p.fill] in GPU_WARP_KEYS and p.group_size > 0) else 1
    group = min(group, n)
    target = max(1, int(33.6e6 // (h * w)))
    chunk = max(group, (target // group) * group)
    return min(chunk, n)


def _depth_geometry(p, depth, n, h, w, resize_depth):
    """Checks the depth batch against the image batch.  Frames of another size are an error for the function
    API (SIG:1586) and are resized on the GPU for the node (GS:141-148, GS:214-220): returns the params to use."""
    if depth.shape[0] != n:   # (callers cut a longer depth batch first: the reference only ever indexes depth_map[i], i < n)
        raise IndexError(f'index {depth.shape[0]} is out of bounds for dimension 0 with size {depth.shape[0]}')
    if tuple(depth.shape[1:3]) == (h, w):
        return p
    if not resize_depth:
        raise AssertionError('Depthmap and the image must have the same size')
    q = CsParams.from_buffer_copy(p)
    q.depth_h, q.depth_w = int(depth.shape[1]), int(depth.shape[2])
    return q


def stereo_batch_device(image, depth, p, out=None, chunk=None, resize_depth=False):
    """The hot path on device-resident tensors.

    image [N,H,W,3] float32 cuda, depth [N,H,W,C] float32 cuda (same H,W; or [N,Hd,Wd,C] with resize_depth=True).
    Returns (stereo, depth_left, depth_right, mask) cuda tensors in the node's layouts.
    Asynchronous on the current stream."""
    if not (image.is_cuda and depth.is_cuda):
        raise ValueError("stereo_batch_device needs CUDA tensors (use stereo_batch_host for CPU tensors)")
    if image.dtype != torch.float32 or depth.dtype != torch.float32:
        raise TypeError("image and depth must be float32")
    image = image.contiguous()
    depth = depth.contiguous()
    n, h, w, ci = image.shape
    if ci != 3:
        raise ValueError("image must have 3 channels")
    if depth.dim() == 3:
        depth = depth.unsqueeze(-1)
    depth = depth[:n]
    p = _depth_geometry(p, depth, n, h, w, resize_depth)
    c = depth.shape[3]
    lib = _lib.lib()
    dev = image.device
    s_shape, d_shape, m_shape = output_shapes(p, n, h, w)
    if out is None:
        out = (torch.empty(s_shape, dtype=torch.float32, device=dev),
               torch.empty(d_shape, dtype=torch.float32, device=dev),
               torch.empty(d_shape, dtype=torch.float32, device=dev),
               torch.empty(m_shape, dtype=torch.float32, device=dev))
    stereo, dl, dr, mask = out
    if chunk is None:
        chunk = default_chunk(p, n, h, w)
    nbytes = lib.cs_workspace_bytes(ctypes.byref(p), int(chunk), h, w)
    ws = _workspace(dev, nbytes)
    with torch.cuda.device(dev):
        _lib.check(lib.cs_stereo_batch(ctypes.byref(p), image.data_ptr(), depth.data_ptr(), n, h, w, c,
                                       stereo.data_ptr(), dl.data_ptr(), dr.data_ptr(), mask.data_ptr(),
                                       ws.data_ptr(), nbytes, _stream_ptr(dev)))
    return stereo, dl, dr, mask


# Outputs are page-locked (so the D2H copies overlap the kernels) only up to this size: a 300-frame 1080p batch is
# 35 GB of results, which should not be pinned wholesale on a workstation.
PIN_LIMIT_BYTES = 4 << 30


def _out_bytes(s_shape, d_shape, m_shape):
    total = 0
    for shp, k in ((s_shape, 1), (d_shape, 2), (m_shape, 1)):
        n = 4
        for v in shp:
            n *= v
        total += k * n
    return total


def stereo_batch_host(image, depth, p, device=0, pin_outputs=True, resize_depth=False, progress=None):
    """The hot path on CPU tensors (what ComfyUI hands the node): chunks are streamed through the
    GPU with upload, kernels and download overlapped inside the library.  Returns CPU tensors.
    progress(frames), if given, is called on this thread each time another chunk of frames is complete (GS:173, GS:262)."""
    if image.is_cuda or depth.is_cuda:
        raise ValueError("stereo_batch_host needs CPU tensors")
    image = image.contiguous().float()
    depth = depth.contiguous().float()
    n, h, w, ci = image.shape
    if ci != 3:
        raise ValueError("image must have 3 channels")
    if depth.dim() == 3:
        depth = depth.unsqueeze(-1)
    depth = depth[:n]
    p = _depth_geometry(p, depth, n, h, w, resize_depth)
    c = depth.shape[3]
    s_shape, d_shape, m_shape = output_shapes(p, n, h, w)
    pin = bool(pin_outputs) and torch.cuda.is_available() and _out_bytes(s_shape, d_shape, m_shape) <= PIN_LIMIT_BYTES
    stereo = torch.empty(s_shape, dtype=torch.float32, pin_memory=pin)
    dl = torch.empty(d_shape, dtype=torch.float32, pin_memory=pin)
    dr = torch.empty(d_shape, dtype=torch.float32, pin_memory=pin)
    mask = torch.empty(m_shape, dtype=torch.float32, pin_memory=pin)
    cb = _lib.PROGRESS_FN(lambda frames, _user: progress(int(frames))) if progress is not None else None
    _lib.check(_lib.lib().cs_stereo_batch_host_progress(ctypes.byref(p), image.data_ptr(), depth.data_ptr(), n, h, w, c,
                                                        stereo.data_ptr(), dl.data_ptr(), dr.data_ptr(), mask.data_ptr(),
                                                        int(device), ctypes.cast(cb, ctypes.c_void_p) if cb else None, None))
    return stereo, dl, dr, mask


# ----------------------------------------------------------------------------------- stage calls (device tensors)
def blur_device(depth, strength, edge_threshold, falloff=1.0, vert_smooth=0, flavor=0):
    """cs_blur on a [n,h,w] float32 CUDA tensor: (left, right).  flavor 1 = the scipy blur of the reference's non-tensor
    branch (SIG:1346-1419), 0 = the torch blur the node runs (SIG:1171-1251)."""
    d = depth.contiguous()
    n, h, w = d.shape
    p = make_params('none', 'left-right', 1.0, depth_blur=True, depth_blur_strength=strength,
                    depth_blur_edge_threshold=edge_threshold, depth_blur_falloff=falloff,
                    depth_blur_vert_smooth=vert_smooth)
    p.blur_flavor = int(flavor)
    left, right = torch.empty_like(d), torch.empty_like(d)
    scratch = torch.empty(2 * n * h * w, dtype=torch.uint8, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().cs_blur(d.data_ptr(), n, h, w, ctypes.byref(p), left.data_ptr(), right.data_ptr(), None,
                                      scratch.data_ptr(), _stream_ptr(d.device)))
    return left, right


def warp_fill_device(image_rgbx, depth, fill_key, divergence, separation, exponent, convergence):
    """cs_warp_fill (apply_stereo_divergence, SIG:1576-1620) for ONE eye: image_rgbx [n,h,w,4] uint8 CUDA, depth [n,h,w]
    float32 as given.  Returns [n,h,w,4] uint8."""
    n, h, w = depth.shape
    out = torch.empty_like(image_rgbx)
    lib = _lib.lib()
    nb = lib.cs_warp_fill_scratch_bytes(n, h, w)
    scratch = torch.empty(nb, dtype=torch.uint8, device=depth.device)
    with torch.cuda.device(depth.device):
        _lib.check(lib.cs_warp_fill(image_rgbx.contiguous().data_ptr(), depth.contiguous().data_ptr(), n, h, w,
                                    FILL_KEYS.index(fill_key), float(divergence), float(separation), float(exponent),
                                    float(convergence), out.data_ptr(), scratch.data_ptr(), nb, _stream_ptr(depth.device)))
    return out


def forward_warp_device(image, depth, div_px, sep_px, exponent, convergence, mesh=False):
    """cs_forward_warp / cs_forward_warp_mesh (forward_warp_gpu SIG:277-450 / forward_warp_mesh SIG:453-689) for ONE eye:
    image [n,h,w,3] float32 CUDA, depth [n,h,w] as given.  Returns (warped [n,h,w,3], mask float32 [n,h,w])."""
    n, h, w = depth.shape
    lib = _lib.lib()
    nb = lib.cs_forward_warp_mesh_scratch_bytes(n, h, w) if mesh else lib.cs_forward_warp_scratch_bytes(n, h, w)
    scratch = torch.empty(nb, dtype=torch.uint8, device=depth.device)
    warped = torch.empty((n, h, w, 3), dtype=torch.float32, device=depth.device)
    mask = torch.empty((n, h, w), dtype=torch.float32, device=depth.device)
    fn = lib.cs_forward_warp_mesh if mesh else lib.cs_forward_warp
    with torch.cuda.device(depth.device):
        _lib.check(fn(image.contiguous().data_ptr(), depth.contiguous().data_ptr(), n, h, w, float(div_px), float(sep_px),
                      float(exponent), float(convergence), warped.data_ptr(), mask.data_ptr(), scratch.data_ptr(), nb,
                      _stream_ptr(depth.device)))
    return warped, mask


def compose_device(left_rgbx, right_rgbx, mode):
    """cs_compose: two [n,h,w,4] uint8 eyes -> (stereo [n,ho,wo,3] float32 = u8 / 255, mask [n,ho,wo])."""
    n, h, w, _ = left_rgbx.shape
    m = MODES.index(mode)
    ho, wo = (h, 2 * w) if m in (0, 1) else ((2 * h, w) if m in (2, 3) else (h, w))
    dev = left_rgbx.device
    stereo = torch.empty((n, ho, wo, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((n, ho, wo), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cs_compose(left_rgbx.contiguous().data_ptr(), right_rgbx.contiguous().data_ptr(), n, h, w, m,
                                         stereo.data_ptr(), mask.data_ptr(), _stream_ptr(dev)))
    return stereo, mask


# ----------------------------------------------------------------------------------- sharding
def shard_range(n, rank, world):
    """Contiguous frame range of `rank`: GPU g gets frames [g*ceil(n/G), ...)."""
    per = -(-n // world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def group_aligned_shard_range(n, rank, world, group):
    """Like shard_range, but cuts only at multiples of `group` so that the GPU-Warp technique's
    sub-batch-wide range tests (quirk Q9) see the same frames as a single-GPU run."""
    ngroups = -(-n // group)
    glo, ghi = shard_range(ngroups, rank, world)
    return min(glo * group, n), min(ghi * group, n)


def stereo_batch_multi_gpu(image, depth, p, devices, resize_depth=False, progress=None):
    """Frame-sharded run over several GPUs of one box from ONE process (used by the node when more
    than one device is visible).  No collective: every device streams its contiguous frame range
    and writes straight into its slice of the (pinned) host outputs, which is in-order assembly by
    construction."""
    n, h, w, _ = image.shape
    if depth.dim() == 3:
        depth = depth.unsqueeze(-1)
    image = image.contiguous().float()
    depth = depth[:n].contiguous().float()
    p = _depth_geometry(p, depth, n, h, w, resize_depth)
    c = depth.shape[3]
    s_shape, d_shape, m_shape = output_shapes(p, n, h, w)
    pin = torch.cuda.is_available() and _out_bytes(s_shape, d_shape, m_shape) <= PIN_LIMIT_BYTES
    outs = (torch.empty(s_shape, dtype=torch.float32, pin_memory=pin),
            torch.empty(d_shape, dtype=torch.float32, pin_memory=pin),
            torch.empty(d_shape, dtype=torch.float32, pin_memory=pin),
            torch.empty(m_shape, dtype=torch.float32, pin_memory=pin))
    group = p.group_size if (FILL_KEYS[p.fill] in GPU_WARP_KEYS and p.group_size > 0) else 1
    lib = _lib.lib()
    errors = []
    done = [0]          # frames finished on any device; the caller's thread forwards it to `progress`
    done_lock = threading.Lock()

    def count(frames, _user):
        with done_lock:
            done[0] += int(frames)

    cb = _lib.PROGRESS_FN(count)

    def work(rank, dev):
        lo, hi = group_aligned_shard_range(n, rank, len(devices), group)
        if hi <= lo:
            return
        try:
            _lib.check(lib.cs_stereo_batch_host_progress(
                ctypes.byref(p), image[lo:hi].data_ptr(), depth[lo:hi].data_ptr(), hi - lo, h, w, c,
                outs[0][lo:hi].data_ptr(), outs[1][lo:hi].data_ptr(), outs[2][lo:hi].data_ptr(),
                outs[3][lo:hi].data_ptr(), int(dev), ctypes.cast(cb, ctypes.c_void_p), None))
        except Exception as e:  # noqa: BLE001 - re-raised on the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r, d)) for r, d in enumerate(devices)]
    for t in threads:
        t.start()
    reported = 0
    while any(t.is_alive() for t in threads):     # the progress bar is only ever touched from the caller's thread
        threads[0].join(timeout=0.02)
        if progress is not None:
            with done_lock:
                now = done[0]
            if now > reported:
                progress(now - reported)
                reported = now
    for t in threads:
        t.join()
    if progress is not None and done[0] > reported:
        progress(done[0] - reported)
    if errors:
        raise errors[0]
    return outs
