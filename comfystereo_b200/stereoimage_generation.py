"""Drop-in mirror of the reference's pipeline-level API (stereoimage_generation.py):

    create_stereoimages(...)      SIG:1422-1574   CPU techniques, one frame, returns PIL images
    create_stereoimages_gpu(...)  SIG:1005-1128   'GPU Warp (Fast)', a sub-batch, returns tensors
    forward_warp_gpu(...)         SIG:277-450     one eye of it, the scatter warp
    forward_warp_mesh(...)        SIG:453-689     one eye of it, the mesh warp the reference uses when moderngl imports

Same names, argument order, defaults, return shapes and exceptions; the work is done by the
sm_100a kernels behind the C ABI (comfystereo_b200/engine.py).  There is no CPU implementation
here: without a B200 and the built library these functions raise.
"""
import os

import torch

from . import engine

# The reference picks the warp of 'GPU Warp (Fast)' from this module flag (SIG:29-36, 1068-1071): forward_warp_mesh when
# moderngl imported, forward_warp_gpu otherwise.  Here nothing depends on OpenGL, so it is a plain switch: False (the
# default) = the scatter warp, bit-exact against the reference; True = the mesh warp, a software rasteriser whose
# coverage/interpolation rule set is fixed and documented because OpenGL's is not.
# COMFYSTEREO_GPU_WARP=mesh sets it at import.
MODERNGL_AVAILABLE = os.environ.get("COMFYSTEREO_GPU_WARP", "scatter").lower() == "mesh"

_CPU_FILLS = ('none', 'naive', 'naive_interpolating', 'polylines_soft', 'polylines_sharp', 'inverse',
              'hybrid_edge', 'none_post', 'inverse_post', 'hybrid_edge_plus')
_ALL_MODES = engine.MODES


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("comfystereo_b200 needs a CUDA (sm_100a) device; it has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_pil_list(t_u8):
    from PIL import Image
    return [Image.fromarray(a) for a in t_u8]


def _float_to_u8(t):
    """The kernels emit u8/255 in float32; x*255 is within 1e-5 of the integer it came from."""
    return torch.round(t * 255.0).clamp_(0, 255).to(torch.uint8)


def create_stereoimages(original_image, depthmap, divergence, separation=0.0, modes=None,
                        stereo_balance=0.0, stereo_offset_exponent=1.0, fill_technique='polylines_sharp',
                        depth_blur_strength=0.0, depth_blur_edge_threshold=6.0,
                        direction_aware_depth_blur=False, return_modified_depth=True, convergence_point=0.5,
                        depth_blur_falloff=1.0, depth_blur_vert_smooth=0):
    """One frame through a CPU technique.  original_image: tensor [3,H,W] (0..1) as the node passes
    it, or an [H,W,3] tensor/array; depthmap: [H,W] (0..1 or 0..255).  Returns
    (list of PIL stereo images, PIL left depth, PIL right depth) when the blur flag is set,
    (list, PIL depth) otherwise, or just the list when return_modified_depth is False."""
    if modes is None:
        modes = ['left-right']
    if not isinstance(modes, list):
        modes = [modes]
    if len(modes) == 0:
        return []
    for mode in modes:
        if mode not in _ALL_MODES:
            raise Exception('Unknown mode')          # SIG:1562
    tensors = isinstance(depthmap, torch.Tensor) and isinstance(original_image, torch.Tensor)
    if not tensors:
        return _create_stereoimages_arrays(original_image, depthmap, divergence, separation, modes, stereo_balance,
                                           stereo_offset_exponent, fill_technique, depth_blur_strength,
                                           depth_blur_edge_threshold, direction_aware_depth_blur, return_modified_depth,
                                           convergence_point, depth_blur_falloff, depth_blur_vert_smooth)
    dev = _device()
    img = original_image
    if img.dim() == 3 and img.shape[0] == 3 and img.shape[2] != 3:
        img = img.permute(1, 2, 0)
    dm = depthmap
    if dm.dim() == 3:
        dm = dm.squeeze()
    h, w = dm.shape
    assert tuple(img.shape[:2]) == (h, w), 'Depthmap and the image must have the same size'
    img = img.to(dev, torch.float32).unsqueeze(0).contiguous()
    dm = dm.to(dev, torch.float32).reshape(1, h, w, 1).contiguous()
    key = fill_technique if fill_technique in _CPU_FILLS else None
    results, left_u8, right_u8 = [], None, None
    for mode in modes:
        if key is None:
            # unknown fill key: apply_stereo_divergence returns the input image (SIG:1620)
            p = engine.make_params('none', mode, 0.0, separation, 0.0, convergence_point,
                                   stereo_offset_exponent, direction_aware_depth_blur, depth_blur_strength,
                                   depth_blur_edge_threshold, depth_blur_falloff, depth_blur_vert_smooth)
        else:
            p = engine.make_params(key, mode, divergence, separation, stereo_balance, convergence_point,
                                   stereo_offset_exponent, direction_aware_depth_blur, depth_blur_strength,
                                   depth_blur_edge_threshold, depth_blur_falloff, depth_blur_vert_smooth)
        stereo, dl, dr, _ = engine.stereo_batch_device(img, dm, p)
        results.append(_float_to_u8(stereo[0]).cpu().numpy())
        if left_u8 is None:
            left_u8 = _float_to_u8(dl[0, :, :, 0]).cpu().numpy()
            right_u8 = _float_to_u8(dr[0, :, :, 0]).cpu().numpy()
    stereo_images = _to_pil_list(results)
    if not return_modified_depth:
        return stereo_images
    from PIL import Image
    if direction_aware_depth_blur:
        return stereo_images, Image.fromarray(left_u8), Image.fromarray(right_u8)
    return stereo_images, Image.fromarray(left_u8)


def _create_stereoimages_arrays(original_image, depthmap, divergence, separation, modes, stereo_balance,
                                stereo_offset_exponent, fill_technique, depth_blur_strength, depth_blur_edge_threshold,
                                direction_aware_depth_blur, return_modified_depth, convergence_point, depth_blur_falloff,
                                depth_blur_vert_smooth):
    """The reference's branch for NON-tensor inputs (numpy arrays, PIL images; SIG:1486-1496, 1520-1526): the image is
    used as it is (uint8), the depth as it is (no x255 rescale), the blur is the scipy one (reflected Sobel, nearest-border
    boxes: engine.blur_device(flavor=1)), and the depth outputs are trunc(clip(depth, 0, 255)).  The node never takes it."""
    import numpy as np
    from PIL import Image
    img = np.asarray(original_image)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise NotImplementedError("non-tensor images must be uint8 [H,W,3] arrays or RGB PIL images")
    depth = np.asarray(depthmap).astype(np.float32)
    if depth.ndim != 2:
        raise NotImplementedError("non-tensor depth maps must be [H,W]")
    dev = _device()
    d = torch.from_numpy(np.ascontiguousarray(depth)).to(dev).unsqueeze(0)
    if direction_aware_depth_blur and depth_blur_strength > 0:
        if int(round(float(depth_blur_strength))) <= 0:
            raise RuntimeError("filter weights array has incorrect shape.")     # what scipy's convolve1d says
        dl, dr = engine.blur_device(d, depth_blur_strength, depth_blur_edge_threshold, depth_blur_falloff,
                                    depth_blur_vert_smooth, flavor=1)
    else:
        dl = dr = d
    rgbx = torch.zeros(img.shape[:2] + (4,), dtype=torch.uint8)
    rgbx[..., :3] = torch.from_numpy(np.ascontiguousarray(img))
    rgbx = rgbx.to(dev).unsqueeze(0)
    ldiv, rdiv = divergence * (1 + stereo_balance), divergence * (1 - stereo_balance)
    known = fill_technique in _CPU_FILLS      # unknown keys: apply_stereo_divergence returns the image, SIG:1620
    if known and (ldiv >= 0.001 or rdiv >= 0.001):
        assert tuple(img.shape[:2]) == tuple(depth.shape), 'Depthmap and the image must have the same size'   # SIG:1586
    left = rgbx if (ldiv < 0.001 or not known) else engine.warp_fill_device(
        rgbx, dl, fill_technique, +1 * ldiv, -1 * separation, stereo_offset_exponent, convergence_point)
    right = rgbx if (rdiv < 0.001 or not known) else engine.warp_fill_device(
        rgbx, dr, fill_technique, -1 * rdiv, separation, stereo_offset_exponent, convergence_point)
    results = []
    for mode in modes:
        stereo, _ = engine.compose_device(left, right, mode)
        results.append(_float_to_u8(stereo[0]).cpu().numpy())
    stereo_images = _to_pil_list(results)
    if not return_modified_depth:
        return stereo_images

    def depth_image(t):
        return Image.fromarray(t[0].clamp(0, 255).to(torch.uint8).cpu().numpy())    # np.clip(...).astype(uint8): truncation

    if direction_aware_depth_blur:
        return stereo_images, depth_image(dl), depth_image(dr)
    return stereo_images, depth_image(d)


def _forward_warp(image_tensor, depth_tensor, divergence_px, separation_px, stereo_offset_exponent, convergence_point, mesh):
    dev = _device()
    img = image_tensor.to(dev, torch.float32).permute(0, 2, 3, 1).contiguous()
    if img.shape[3] != 3:
        raise NotImplementedError("the warp kernels take 3-channel images")
    dep = depth_tensor.to(dev, torch.float32).contiguous()
    warped, mask = engine.forward_warp_device(img, dep, divergence_px, separation_px, stereo_offset_exponent,
                                              convergence_point, mesh=mesh)
    return warped.permute(0, 3, 1, 2), mask > 0.5


def forward_warp_gpu(image_tensor, depth_tensor, divergence_px, separation_px, stereo_offset_exponent,
                     convergence_point=0.5, max_stretch=8):
    """SIG:277-450.  image_tensor [B,3,H,W] (0..1), depth_tensor [B,H,W] (0..1 or 0..255; /255 when ANY frame's max > 1).
    Returns (warped [B,3,H,W], unfilled bool [B,H,W]) on the CUDA device.  max_stretch is the reference's scatter round
    count; only its default 8 is implemented."""
    if max_stretch != 8:
        raise NotImplementedError("forward_warp_gpu: max_stretch is fixed at the reference's default 8")
    return _forward_warp(image_tensor, depth_tensor, divergence_px, separation_px, stereo_offset_exponent,
                         convergence_point, False)


def forward_warp_mesh(image_tensor, depth_tensor, divergence_px, separation_px, stereo_offset_exponent,
                      convergence_point=0.5, gradient_threshold=1.5, max_stretch=8):
    """SIG:453-689, rasterised in software with the rule set of DESIGN.md section 9.  Same shapes as forward_warp_gpu;
    the mask is the pre-fill gap map.  gradient_threshold is fixed at the reference's default 1.5 (max_stretch is unused
    there too)."""
    if gradient_threshold != 1.5:
        raise NotImplementedError("forward_warp_mesh: gradient_threshold is fixed at the reference's default 1.5")
    return _forward_warp(image_tensor, depth_tensor, divergence_px, separation_px, stereo_offset_exponent,
                         convergence_point, True)


def create_stereoimages_gpu(image_tensor, depth_tensor, divergence, separation=0.0, modes=None,
                            stereo_balance=0.0, stereo_offset_exponent=1.0, convergence_point=0.5,
                            depth_blur_strength=0.0, depth_blur_edge_threshold=6.0,
                            direction_aware_depth_blur=False, depth_blur_falloff=1.0,
                            depth_blur_vert_smooth=0):
    """A sub-batch through 'GPU Warp (Fast)' (the scatter warp, or the mesh warp when MODERNGL_AVAILABLE is set).
    image_tensor [B,3,H,W], depth_tensor [B,H,W].
    Returns (list of [B,3,Ho,Wo] tensors, left_depth [B,H,W], right_depth [B,H,W], mask bool [B,H,W])
    on the CUDA device, like the reference does when CUDA is available."""
    if modes is None:
        modes = ['left-right']
    if not isinstance(modes, list):
        modes = [modes]
    if len(modes) == 0:
        return [], None, None, None
    for mode in modes:
        if mode not in _ALL_MODES:
            raise ValueError(f'Unknown mode: {mode}')  # SIG:1120
    dev = _device()
    b, _, h, w = image_tensor.shape
    img = image_tensor.to(dev, torch.float32).permute(0, 2, 3, 1).contiguous()
    dm = depth_tensor.to(dev, torch.float32).reshape(b, h, w, 1).contiguous()
    results, left, right, mask = [], None, None, None
    for mode in modes:
        p = engine.make_params('gpu_warp_mesh' if MODERNGL_AVAILABLE else 'gpu_warp', mode, divergence, separation, stereo_balance, convergence_point,
                               stereo_offset_exponent, direction_aware_depth_blur, depth_blur_strength,
                               depth_blur_edge_threshold, depth_blur_falloff, depth_blur_vert_smooth,
                               group_size=b)
        stereo, dl, dr, m = engine.stereo_batch_device(img, dm, p, chunk=b)
        results.append(stereo.permute(0, 3, 1, 2))
        if left is None:
            # the node-level depth outputs are clamped copies (GS:163-166); at function level the
            # reference returns them unclamped, which only differs for depth outside [0, 255]
            left, right, mask = dl[..., 0], dr[..., 0], m > 0.5
    return results, left, right, mask
