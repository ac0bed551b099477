import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from comfystereo_b200 import engine, synthetic as syn
dev = torch.device('cuda', 0)
for name, (lh, lw, lfill, lblur) in {"c0_512_naive": (512, 512, "naive", False), "c1_1080p_poly": (1080, 1920, "polylines_sharp", True)}.items():
    li = torch.from_numpy(syn.make_image(1, lh, lw, seed=7)).to(dev); ld = torch.from_numpy(syn.make_depth(1, lh, lw, "scene", seed=7)).to(dev)
    lp = engine.make_params(lfill, "left-right", 3.5, 0.0, 0.0, 0.5, 2.0, lblur, 20.0, 20.0, 2.0, 6)
    lo = engine.stereo_batch_device(li, ld, lp)
    ts = []
    for _ in range(50):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); engine.stereo_batch_device(li, ld, lp, out=lo); a1.record(); a1.synchronize()
        ts.append(a0.elapsed_time(a1) * 1e3)
    print(name, "median %.1f us  min %.1f us" % (np.median(ts), np.min(ts)))
