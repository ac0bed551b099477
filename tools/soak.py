"""Dev soak: a long 1080p batch through the node (host tensors), Hybrid Edge like BASELINE config 3, checks a few frames."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from comfystereo_b200 import StereoImageNode, synthetic as syn
import oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
fill = sys.argv[2] if len(sys.argv) > 2 else "Imperfect fill - Hybrid Edge"
base_i = syn.make_image(4, 1080, 1920, seed=3); base_d = syn.make_depth(4, 1080, 1920, "scene", seed=3)
img = torch.from_numpy(np.tile(base_i, (n // 4 + 1, 1, 1, 1))[:n]); dep = torch.from_numpy(np.tile(base_d, (n // 4 + 1, 1, 1, 1))[:n])
p = dict(divergence=3.5, separation=0.0, modes="left-right", stereo_balance=0.0, convergence_point=0.5, stereo_offset_exponent=2.0,
         fill_technique=fill, depth_blur_edge_threshold=20.0, depth_blur_strength=20.0, depth_map_blur=True,
         depth_blur_falloff=2.0, depth_blur_vert_smooth=6, batch_size=12)
node = StereoImageNode()
node.generate(img[:4], dep[:4], **p)
t0 = time.perf_counter(); out = node.generate(img, dep, **p); dt = time.perf_counter() - t0
print(f"{n} frames {fill}: {dt:.2f} s = {n/dt:.1f} fps end to end (pageable inputs), pinned outputs: {out[0].is_pinned()}")
want = orc.node_generate(base_i[:1], base_d[:1], **p)
for k in (0, 4 * ((n - 1) // 4)):
    d = np.abs(np.rint(out[0][k].numpy() * 255) - np.rint(want[0][0] * 255)).max()
    assert d <= 1, d
    assert np.array_equal(np.rint(out[1][k].numpy() * 255), np.rint(want[1][0] * 255))
print("frames 0 and", 4 * ((n - 1) // 4), "match the oracle")
