"""Development aid: checks the polylines kernel's per-column logic on the CPU (tools/poly_model.cpp, which compiles the
kernel's own csrc/cs_poly_core.cuh for the host) against the oracle:
  * exact_column == oracle bit for bit (unless a list replay gave up: those rows go to the sequential kernel);
  * every column the float32 path certifies == exact_column.
Prints how many columns needed the exact path.   python tools/poly_model_check.py [quick]"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402
from comfystereo_b200 import synthetic as syn  # noqa: E402

SO = "/tmp/libpoly_model.so"
subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO,
                       os.path.join(ROOT, "tools", "poly_model.cpp")])
lib = ctypes.CDLL(SO)


def model(img, nd, div_px, sep_px, expo, sharp, stats):
    h, w = nd.shape
    oe, of, ok = np.zeros((h, w, 3), np.uint8), np.zeros((h, w, 3), np.uint8), np.zeros((h, w), np.uint8)
    gave = np.zeros(h, np.int32)
    for y in range(h):
        gave[y] = lib.poly_model_row(img[y].ctypes.data_as(ctypes.c_void_p), nd[y].ctypes.data_as(ctypes.c_void_p), w,
                                     ctypes.c_double(div_px), ctypes.c_double(sep_px), ctypes.c_double(expo), int(sharp),
                                     oe[y].ctypes.data_as(ctypes.c_void_p), of[y].ctypes.data_as(ctypes.c_void_p),
                                     ok[y].ctypes.data_as(ctypes.c_void_p), stats.ctypes.data_as(ctypes.c_void_p))
    return oe, of, ok.astype(bool), gave.astype(bool)


def check(name, img, nd, div_px, sep_px, expo, sharp):
    img = np.ascontiguousarray(img, np.uint8)
    nd = np.ascontiguousarray(nd, np.float32)
    stats = np.zeros(8, np.int32)
    ref = orc.polylines(img, nd, div_px, sep_px, expo, sharp)
    oe, of, ok, gave = model(img, nd, div_px, sep_px, expo, sharp, stats)
    good_rows = ~gave
    bad_exact = int(np.any(oe[good_rows] != ref[good_rows], axis=-1).sum())
    bad_fast = int((np.any(of != ref, axis=-1) & ok).sum())
    n = ok.size
    print(f"{name:46s} sharp={int(sharp)} exact!=oracle {bad_exact:5d}  certified!=oracle {bad_fast:5d}  "
          f"uncertified {100.0 * (1 - ok.mean()):6.3f}%  hard {100.0 * stats[0] / max(stats[7], 1):5.2f}%  "
          f"codes0-3 {stats[1]}/{stats[2]}/{stats[3]}/{stats[4]}  gave_up rows {int(gave.sum())}/{len(gave)}")
    return bad_exact + bad_fast


def main():
    quick = len(sys.argv) > 1
    bad = 0
    h, w = 1080, 1920
    rows = slice(0, h, 40 if quick else 9)
    img = (syn.make_image(1, h, w, seed=0)[0] * 255).astype(np.uint8)
    dep = syn.make_depth(1, h, w, "scene", seed=0)[0, ..., 0] * np.float32(255)
    L, R = orc.blur(dep, 20.0, 20.0, 2.0, 6)
    for sharp in (True, False):
        for eye, (d, div) in enumerate(((L, 3.5), (R, -3.5))):
            nd = orc.normalize(d, 0.5)
            bad += check(f"bench scene eye {eye}", img[rows], nd[rows], div / 100 * w, 0.0, 2.0, sharp)
    for kind in ("noise", "quant", "steps", "card", "flat"):
        hh, ww = 64, 640
        im = (syn.make_image(1, hh, ww, seed=3)[0] * 255).astype(np.uint8)
        d = syn.make_depth(1, hh, ww, kind, seed=1)[0, ..., 0] * np.float32(255)
        for conv, expo, divp, sep in ((0.5, 2.0, 3.5, 0.0), (0.0, 1.0, -6.0, 1.0), (1.0, 0.7, 10.0, -2.5), (0.3, 2.0, 15.0, 0.0)):
            nd = orc.normalize(d, conv)
            for sharp in (True, False):
                bad += check(f"{kind} conv {conv} expo {expo} div {divp} sep {sep}", im, nd, divp / 100 * ww, sep / 100 * ww, expo, sharp)
    rng = np.random.default_rng(7)
    for i in range(20 if quick else 150):
        im, d, div, sep, expo, conv = syn.fuzz_case(rng)
        if d.shape[1] < 2:
            continue
        nd = orc.normalize(d, conv)
        wv = d.shape[1]
        for sharp in (True, False):
            bad += check(f"fuzz {i}", im, nd, div / 100 * wv, sep / 100 * wv, expo, sharp)
    print("TOTAL MISMATCHES", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
