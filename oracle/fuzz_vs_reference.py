"""Fuzz the oracle against the UNMODIFIED reference (needs /root/reference; build container only).

TEST INFRASTRUCTURE ONLY.  The committed fixtures pin the oracle on a fixed set of inputs; this script hunts for
rare disagreements (rounding cases, dark images, extreme parameters) on random ones:

    python oracle/fuzz_vs_reference.py [iterations] [seed]

Every disagreement is printed with the parameters that reproduce it; anything found becomes a fixture
(oracle/make_golden.py) after the oracle is fixed.
"""
import sys
import os
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
import oracle  # noqa: E402
from comfystereo_b200 import synthetic as syn  # noqa: E402

FILLS = ['none', 'naive', 'naive_interpolating', 'polylines_soft', 'polylines_sharp', 'inverse', 'hybrid_edge',
         'none_post', 'inverse_post', 'hybrid_edge_plus']


random_case = syn.fuzz_case


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    assert ref_loader.reference_available(), "needs /root/reference"
    Node = ref_loader.load_node_class()
    sig = sys.modules[Node.__module__].sig
    rng = np.random.default_rng(seed)
    bad = {}
    for it in range(iters):
        img, d, div, sep, expo, conv = random_case(rng)
        for fill in FILLS:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = np.asarray(sig.apply_stereo_divergence(img.copy(), d.copy(), div, sep, expo, fill, conv))
            got = oracle.apply_stereo_divergence(img, d, div, sep, expo, fill, conv)
            diff = np.abs(ref.astype(np.int32) - got.astype(np.int32))
            if diff.max() > 0:
                bad[fill] = bad.get(fill, 0) + 1
                print(f"MISMATCH it={it} seed={seed} fill={fill} shape={img.shape[:2]} div={div} sep={sep} expo={expo} "
                      f"conv={conv}: {int((diff > 0).sum())} values, max {int(diff.max())}")
    print("iterations", iters, "mismatching cases per fill:", bad if bad else "none")
    # forward_warp_gpu (SIG:277-450) on torch-CPU: mask bit-exact, image to float32 rounding
    import torch
    gbad = 0
    for it in range(max(1, iters // 4)):
        img, d, div, sep, expo, conv = random_case(rng)
        h, w = d.shape
        imgf = (img.astype(np.float32) / np.float32(255)).transpose(2, 0, 1).copy()
        d01 = (d / np.float32(255)).astype(np.float32)
        div_px, sep_px = div / 100.0 * w, sep / 100.0 * w
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rw, rm = sig.forward_warp_gpu(torch.from_numpy(imgf)[None], torch.from_numpy(d01)[None], div_px, sep_px,
                                          expo, conv)
        ow, om = oracle.gpuwarp_eye(imgf, d01, div_px, sep_px, expo, conv)
        mm = int((rm[0].numpy().astype(bool) != om.astype(bool)).sum())
        err = float(np.abs(rw[0].numpy() - ow).max())
        tol = 2e-5 if expo in (1.0, 2.0) else 1e-4
        if mm or err > tol:
            gbad += 1
            print(f"GPUWARP MISMATCH it={it} seed={seed} shape={(h, w)} div_px={div_px} sep_px={sep_px} expo={expo} "
                  f"conv={conv}: mask {mm}, image err {err:.3g}")
    print("forward_warp_gpu cases", max(1, iters // 4), "mismatching:", gbad)
    # the node glue end to end with the blur off (so that everything is exact): modes, balance, pass-through eyes,
    # masks, depth outputs, sub-batches, depth given as 1 / 3 channels, on the 0..1 or the 0..255 scale, other sizes
    names = list(oracle.FILL_NAME_TO_KEY)[:8]
    nbad = 0
    ncases = max(1, iters // 4)
    for it in range(ncases):
        n = int(rng.integers(1, 4))
        h, w = int(rng.integers(2, 14)), int(rng.integers(4, 90))
        img = rng.random((n, h, w, 3), dtype=np.float32)
        if it % 5 == 0:
            img[:, :, : w // 3] = 0.0
        ch = int(rng.choice([1, 3]))
        dh, dw = (h, w) if it % 3 else (int(rng.integers(2, 20)), int(rng.integers(2, 60)))
        dep = rng.random((n, dh, dw, 1), dtype=np.float32) * np.float32(rng.choice([1.0, 255.0]))
        dep = np.repeat(dep, ch, axis=-1)
        if dh != h or dw != w:
            os.environ["ATEN_CPU_CAPABILITY"] = "default"   # only read at torch start-up: see the resize note in DESIGN.md
        kw = dict(divergence=float(rng.choice([0.05, 2.0, 4.5, 9.0, 15.0])), separation=float(rng.choice([0.0, 1.5, -3.0])),
                  modes=str(rng.choice(oracle.MODES[:5])), stereo_balance=float(rng.choice([0.0, 0.5, -0.95, 0.95])),
                  convergence_point=float(rng.choice([0.0, 0.5, 1.0])), stereo_offset_exponent=float(rng.choice([1.0, 2.0])),
                  fill_technique=str(rng.choice(names)), depth_blur_edge_threshold=20.0, depth_blur_strength=20.0,
                  depth_map_blur=False, depth_blur_falloff=2.0, depth_blur_vert_smooth=6, batch_size=int(rng.integers(1, 4)))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = [o.numpy() for o in Node().generate(torch.from_numpy(img), torch.from_numpy(dep), **kw)]
        got = oracle.node_generate(img, dep, **kw)
        gw = kw["fill_technique"] == 'GPU Warp (Fast)'
        resized = (dh, dw) != (h, w)
        errs = [float(np.abs(a - b).max()) if a.shape == b.shape else float("inf") for a, b in zip(ref, got)]
        # a resized depth goes through torch's (possibly FMA-contracting) bilinear kernel: not bit-comparable here
        tol = [2e-5 if gw else 0.0, 1e-6 if gw else 0.0, 1e-6 if gw else 0.0, 0.0]
        if resized:
            continue_ok = all(a.shape == b.shape for a, b in zip(ref, got))
            if not continue_ok:
                nbad += 1
                print(f"NODE SHAPE MISMATCH it={it} seed={seed} {kw}")
            continue
        if any(e > t for e, t in zip(errs, tol)):
            nbad += 1
            print(f"NODE MISMATCH it={it} seed={seed} n={n} h={h} w={w} ch={ch} errs={errs} {kw}")
    print("node cases", ncases, "mismatching:", nbad)
    return 1 if (bad or gbad or nbad) else 0


if __name__ == "__main__":
    sys.exit(main())
