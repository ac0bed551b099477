"""ComfyUI node "Stereo Image Node" -- the drop-in boundary (reference: GenerateStereo.py, GS:46-361).

Identical widget schema (names, order, defaults, ranges), RETURN_TYPES / RETURN_NAMES / FUNCTION,
`generate` signature and output conventions (four float32 CPU tensors: stereoscope [N,Ho,Wo,3],
blurred_depthmap_left/right [N,H,W,3], no_fill_imperfect_mask [N,Hm,Wm]).  Inside, the whole batch is
handed to the sm_100a library in one call: frames stream host -> GPU -> host in overlapped chunks
(and are sharded frame-wise over the visible GPUs when COMFYSTEREO_MULTI_GPU is set).

"GPU Warp (Fast)" is the reference's forward_warp_gpu (SIG:277-450) by default -- what the reference runs when moderngl
is not importable -- and bit-exact against it.  With moderngl installed the reference switches to an OpenGL mesh
rasteriser (forward_warp_mesh, SIG:1067-1071) whose fragments depend on the GL implementation; COMFYSTEREO_GPU_WARP=mesh
(stereoimage_generation.MODERNGL_AVAILABLE) selects this package's software rasteriser for that mesh instead.
"""
import os

import torch
import torch.distributed

from . import engine, stereoimage_generation

try:  # inside ComfyUI
    from comfy.utils import ProgressBar
except Exception:  # noqa: BLE001 - stand-alone use (tests, bench): same surface, no UI
    class ProgressBar:
        def __init__(self, total):
            self.total, self.current = total, 0

        def update(self, value):
            self.current += value

MODES = ["left-right", "right-left", "top-bottom", "bottom-top", "red-cyan-anaglyph"]
FILL_TECHNIQUES = ['GPU Warp (Fast)', 'No fill', 'No fill - Reverse projection', 'Imperfect fill - Hybrid Edge',
                   'Fill - Naive', 'Fill - Naive interpolating', 'Fill - Polylines Soft', 'Fill - Polylines Sharp']


def _gray_depth(depth_map, gpu_branch):
    """N1: depth IMAGE [N,H,W,C] -> what the reference feeds the pipeline.  3 channels are weighted
    on the device by the library (it takes C = 3 directly); 1 channel is used as is.  Any other channel
    count: the GPU-Warp branch takes channel 0 (GS:138-139); the per-frame branch of the reference
    would fail on it, so it is rejected here."""
    c = depth_map.shape[3]
    if c in (1, 3):
        return depth_map
    if gpu_branch:
        return depth_map[..., :1]
    raise ValueError(f"depth_map with {c} channels is not supported by this fill technique")


class StereoImageNode:
    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "image": ("IMAGE",),
                "depth_map": ("IMAGE",),
                "modes": (list(MODES),),
                "fill_technique": (list(FILL_TECHNIQUES), {
                    "default": "GPU Warp (Fast)",
                    "tooltip": "How disoccluded areas are treated. All techniques run as B200 CUDA kernels. "
                               "GPU Warp (Fast) = the reference's scatter warp (forward_warp_gpu); COMFYSTEREO_GPU_WARP=mesh switches to the mesh warp (forward_warp_mesh)."}),
            },
            "optional": {
                "divergence": ("FLOAT", {"default": 4.5, "min": 0.05, "max": 15, "step": 0.01,
                                         "tooltip": "Strength of the 3D effect, percent of image width."}),
                "separation": ("FLOAT", {"default": 0, "min": -5, "max": 5, "step": 0.01,
                                         "tooltip": "Uniform horizontal offset between the eyes, percent of width."}),
                "stereo_balance": ("FLOAT", {"default": 0, "min": -0.95, "max": 0.95, "step": 0.05,
                                             "tooltip": "Shifts the divergence towards the left (+) or right (-) eye."}),
                "convergence_point": ("FLOAT", {"default": 0.5, "min": 0.0, "max": 1.0, "step": 0.05,
                                                "tooltip": "Normalised depth that stays on the screen plane."}),
                "stereo_offset_exponent": ("FLOAT", {"default": 2, "min": 0.1, "max": 2, "step": 0.1,
                                                     "tooltip": "Exponent of the depth-to-disparity curve."}),
                "depth_map_blur": ("BOOLEAN", {"default": True,
                                               "tooltip": "Edge-aware directional blur of the depth map before warping."}),
                "depth_blur_edge_threshold": ("FLOAT", {"default": 20, "min": 0.1, "max": 60, "step": 0.1,
                                                        "tooltip": "Gradient threshold that marks a depth edge."}),
                "depth_blur_strength": ("FLOAT", {"default": 20, "min": 0.1, "max": 200, "step": 0.1,
                                                  "tooltip": "Width of the blur and of its influence around edges (px)."}),
                "depth_blur_falloff": ("FLOAT", {"default": 2.0, "min": 0.1, "max": 4.0, "step": 0.1,
                                                 "tooltip": "Exponent of the blur weight decay away from an edge."}),
                "depth_blur_vert_smooth": ("INT", {"default": 6, "min": 0, "max": 15, "step": 1,
                                                   "tooltip": "Vertical smoothing radius of the blur weights (px)."}),
                "batch_size": ("INT", {"default": 12, "min": 1, "max": 64, "step": 1,
                                       "tooltip": "GPU Warp: frames per sub-batch (range tests are sub-batch wide)."}),
            }
        }

    RETURN_TYPES = ("IMAGE", "IMAGE", "IMAGE", "MASK")
    RETURN_NAMES = ("stereoscope", "blurred_depthmap_left", "blurred_depthmap_right", "no_fill_imperfect_mask")
    FUNCTION = "generate"

    def generate(self, image, depth_map, divergence, separation, modes,
                 stereo_balance, convergence_point, stereo_offset_exponent, fill_technique,
                 depth_blur_edge_threshold, depth_blur_strength, depth_map_blur, depth_blur_falloff=1.0,
                 depth_blur_vert_smooth=0, batch_size=4):
        key = engine.FILL_NAME_TO_KEY.get(fill_technique, 'gpu_warp')   # unknown names -> GPU Warp, GS:102
        gpu_branch = key == 'gpu_warp'
        if gpu_branch and stereoimage_generation.MODERNGL_AVAILABLE:     # SIG:1068-1071
            key = 'gpu_warp_mesh'
        total = len(image)
        pbar = ProgressBar(total)
        image = image.float() if image.dtype != torch.float32 else image
        depth_map = depth_map.float() if depth_map.dtype != torch.float32 else depth_map
        if depth_map.dim() == 3:
            depth_map = depth_map.unsqueeze(-1)
        # a depth batch of another size is resized on the GPU (gray first, then bilinear: GS:141-148, GS:214-220)
        depth = _gray_depth(depth_map, gpu_branch)
        group = min(int(batch_size), total) if gpu_branch else 0        # GS:119
        p = engine.make_params(key, modes, divergence, separation, stereo_balance, convergence_point,
                               stereo_offset_exponent, depth_map_blur, depth_blur_strength,
                               depth_blur_edge_threshold, depth_blur_falloff, depth_blur_vert_smooth,
                               group_size=group)
        if image.is_cuda:
            outs = engine.stereo_batch_device(image, depth.to(image.device), p, resize_depth=True)
            outs = tuple(o.cpu() for o in outs)
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("comfystereo_b200 needs a CUDA (sm_100a) device; it has no CPU fallback")
            # One GPU by default, like the reference.  COMFYSTEREO_MULTI_GPU=1 (or =<count>) lets this process shard the
            # batch frame-wise over the visible GPUs -- opt-in, because every device then holds its own workspace and
            # page-locked staging buffers (a shared server may not want that).  Never inside a torch.distributed job,
            # where each rank owns one GPU already.
            want = os.environ.get("COMFYSTEREO_MULTI_GPU", "0")
            in_job = torch.distributed.is_available() and torch.distributed.is_initialized()
            ndev = 1
            if want not in ("", "0") and not in_job and os.environ.get("COMFYSTEREO_SINGLE_DEVICE", "0") != "1":
                ndev = torch.cuda.device_count() if want == "1" else min(int(want), torch.cuda.device_count())
            # progress per chunk of frames, as the reference reports it per frame / sub-batch (GS:173, GS:262)
            if ndev > 1 and total >= 2 * ndev:
                outs = engine.stereo_batch_multi_gpu(image, depth, p, list(range(ndev)), resize_depth=True, progress=pbar.update)
            else:
                outs = engine.stereo_batch_host(image, depth, p, device=torch.cuda.current_device(), resize_depth=True,
                                                progress=pbar.update)
            return outs
        pbar.update(total)
        return outs


NODE_CLASS_MAPPINGS = {"StereoImageNode": StereoImageNode}
NODE_DISPLAY_NAME_MAPPINGS = {"StereoImageNode": "Stereo Image Node"}
