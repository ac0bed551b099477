"""Dev: end-to-end cost of the host path for pageable vs pinned inputs / outputs (1080p Naive, so kernels are cheap)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from comfystereo_b200 import engine, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
base_i = syn.make_image(4, 1080, 1920, seed=3); base_d = syn.make_depth(4, 1080, 1920, "scene", seed=3)
img = torch.from_numpy(np.tile(base_i, (n // 4 + 1, 1, 1, 1))[:n]); dep = torch.from_numpy(np.tile(base_d, (n // 4 + 1, 1, 1, 1))[:n])
p = engine.make_params("naive", "left-right", 3.5, 0, 0, 0.5, 2.0, True, 20, 20, 2.0, 6)
imgp, depp = img.pin_memory(), dep.pin_memory()
for name, a, b, pin in (("pinned in, pinned out", imgp, depp, True), ("pageable in, pinned out", img, dep, True),
                        ("pinned in, pageable out", imgp, depp, False), ("pageable in, pageable out", img, dep, False)):
    engine.stereo_batch_host(a[:4], b[:4], p, pin_outputs=pin)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); o = engine.stereo_batch_host(a, b, p, pin_outputs=pin); ts.append(time.perf_counter() - t0); del o
    print(f"{name:28s} {n} frames: best {min(ts)*1e3:8.1f} ms = {n/min(ts):7.1f} fps   (all: {[round(t*1e3) for t in ts]})")
