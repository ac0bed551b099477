"""Lean device-resident timing of one configuration (development aid; bench.py is the contract).
   python tools/kbench.py [--fill F] [--frames N] [--steps K] [--width W --height H] [--divergence D] [--balance B] [--mode M]
Prints frames/s and the per-kernel CUDA-event times per step."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from comfystereo_b200 import _lib, engine, synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fill", default="polylines_sharp")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--divergence", type=float, default=3.5)
    ap.add_argument("--balance", type=float, default=0.0)
    ap.add_argument("--separation", type=float, default=0.0)
    ap.add_argument("--mode", default="left-right")
    ap.add_argument("--no-blur", action="store_true")
    ap.add_argument("--kind", default="scene")
    ap.add_argument("--tag", default="")
    ap.add_argument("--chunk", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.lib()
    n, h, w = a.frames, a.height, a.width
    base = min(n, 4)
    img = np.tile(syn.make_image(base, h, w, seed=0), (-(-n // base), 1, 1, 1))[:n]
    dep = np.tile(syn.make_depth(base, h, w, a.kind, seed=0), (-(-n // base), 1, 1, 1))[:n]
    group = min(12, n) if a.fill.startswith("gpu_warp") else 0
    p = engine.make_params(a.fill, a.mode, a.divergence, a.separation, a.balance, 0.5, 2.0, not a.no_blur, 20.0, 20.0, 2.0, 6,
                           group_size=group)
    img_d, dep_d = torch.from_numpy(img).to(dev), torch.from_numpy(dep).to(dev)
    outs = engine.stereo_batch_device(img_d, dep_d, p)
    for _ in range(3):
        engine.stereo_batch_device(img_d, dep_d, p, out=outs, chunk=(a.chunk or None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        engine.stereo_batch_device(img_d, dep_d, p, out=outs, chunk=(a.chunk or None))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    lib.cs_profile_enable(1)
    for _ in range(a.steps):
        engine.stereo_batch_device(img_d, dep_d, p, out=outs, chunk=(a.chunk or None))
    torch.cuda.synchronize()
    lib.cs_profile_enable(0)
    nk = lib.cs_profile_kernel_count()
    k_ms = (ctypes.c_double * nk)()
    k_n = (ctypes.c_longlong * nk)()
    lib.cs_profile_collect(k_ms, k_n)
    ks = {lib.cs_profile_kernel_name(i).decode(): round(k_ms[i] / a.steps, 3) for i in range(nk) if k_n[i] > 0}
    if hasattr(lib, "cs_poly_ticks") or True:
        try:
            tk = (ctypes.c_ulonglong * 16)()
            lib.cs_poly_ticks(tk, 1)
            tot = sum(tk[:11]) or 1
            print("   poly ticks share:", " ".join(f"{i}:{100.0 * tk[i] / tot:.1f}%" for i in range(11)), f"(total {tot / 1e9:.3f} Gcycles of thread 0; {tk[13]} CTAs, {tot / max(tk[13], 1):.0f} cycles each, {tk[12] / max(tk[13], 1):.2f} uncertified columns per CTA)")
        except AttributeError:
            pass
    print(f"{a.tag} {a.fill} {w}x{h} x{n}: {n / ms * 1e3:9.1f} fps  {ms:7.3f} ms/step  {ks}", flush=True)


if __name__ == "__main__":
    main()
