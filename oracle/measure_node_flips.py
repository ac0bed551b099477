"""Measures, per node-level fixture with the blur on, how many output pixels of THIS implementation (own blur) differ
from the reference's own output, and writes tests/golden/node_flip_counts.json.  CPU only: the oracle's uint8 outputs
equal the CUDA path's bit for bit (tests/test_gpu_parity.py), so the counts hold for both.

Why pixels differ at all: torch's conv2d summation order is unspecified, so the blurred depth differs from the
reference's by float32 rounding (<= 2e-4 on the 0..255 scale, pinned by the blur stage tests); where that noise moves a
shift across an integer, or decides between equal plateau levels, a different source pixel is picked.  With the
reference's blurred depth injected everything downstream is bit-exact (test_node_vs_reference), so the blurred depth is
the only cause; the tests hold every fixture to the counts recorded here instead of a loose common bound.

    python oracle/measure_node_flips.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402
from conftest import load_manifest, load_golden  # noqa: E402
from comfystereo_b200 import synthetic as syn  # noqa: E402


def flips(spec, g, stereo, mask):
    """(pixels off by more than 1 LSB, mask pixels that differ) against the reference fixture."""
    if spec["params"]["fill_technique"] == 'GPU Warp (Fast)':
        px = int((np.abs(stereo - g["stereo"]).max(axis=-1) > 1.0 / 255).sum())
        mk = int(((mask > 0).astype(np.uint8) != g["mask"]).sum())
    else:
        q = lambda a: np.rint(a * 255.0).astype(np.uint8)
        px = int((np.abs(q(stereo).astype(np.int32) - g["stereo"].astype(np.int32)).max(axis=-1) > 1).sum())
        mk = int((q(mask) != g["mask"]).sum())
    return px, mk


def main():
    out = {}
    for spec in load_manifest()["node"]:
        g = load_golden("node", spec["name"])
        if "blur_l" not in g.files:
            continue
        img = syn.make_image(spec["n"], spec["h"], spec["w"], seed=spec["seed"], black_box=spec["black_box"])
        dep = syn.make_depth(spec["n"], spec.get("dh", spec["h"]), spec.get("dw", spec["w"]), spec["kind"],
                             seed=spec["seed"], channels=spec["channels"], scale255=spec["scale255"])
        stereo, dl, dr, mask = orc.node_generate(img, dep, **spec["params"])
        px, mk = flips(spec, g, stereo, mask)
        total = int(np.prod(stereo.shape[:-1]))
        out[spec["name"]] = {"pixels": px, "mask": mk, "of": total, "kind": spec["kind"]}
        print(f"{spec['name']:40s} {spec['kind']:6s} {spec['params']['fill_technique']:32s} px {px:7d} mask {mk:7d} of {total}")
    with open(os.path.join(ROOT, "tests", "golden", "node_flip_counts.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
