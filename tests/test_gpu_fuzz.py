"""-m gpu: random small cases through every fill, CUDA (C ABI) against the CPU oracle.  The same generator drives
oracle/fuzz_vs_reference.py, which holds the oracle to the unmodified reference in the build container; together they
chain CUDA == oracle == reference on inputs no fixture was written for (dark / black / bright images, noisy / stepped /
flat depth, negative divergence, odd sizes)."""
import numpy as np
import pytest
import torch

from comfystereo_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

FILLS = ['none', 'naive', 'naive_interpolating', 'polylines_soft', 'polylines_sharp', 'inverse', 'hybrid_edge',
         'none_post', 'inverse_post', 'hybrid_edge_plus']


@pytest.fixture(scope="module")
def gu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("seed", range(16))
def test_random_stage_cases(gu, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        img, d, div, sep, expo, conv = syn.fuzz_case(rng)
        for fill in FILLS:
            got = gu.warp_fill(img, d, fill, div, sep, expo, conv)[..., :3]
            want = oracle.apply_stereo_divergence(img, d, div, sep, expo, fill, conv)
            diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
            if fill.startswith('hybrid'):
                assert diff.max() <= 1, (fill, img.shape, div, sep, expo, conv, int(diff.max()))
            elif expo in (1.0, 2.0):
                assert diff.max() == 0, (fill, img.shape, div, sep, expo, conv, int((diff > 0).sum()))
            else:   # CUDA pow vs libm pow: an index can flip only within ~1e-16 of an integer
                assert (diff > 0).mean() <= 1e-3, (fill, img.shape, div, sep, expo, conv, int((diff > 0).sum()))


@pytest.mark.parametrize("seed", range(6))
def test_random_forward_warp_cases(gu, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(10):
        img, d, div, sep, expo, conv = syn.fuzz_case(rng)
        h, w = d.shape
        imgf = (img.astype(np.float32) / np.float32(255))
        d01 = (d / np.float32(255)).astype(np.float32)
        div_px, sep_px = div / 100.0 * w, sep / 100.0 * w
        warped, mask = gu.forward_warp(imgf[None], d01[None], div_px, sep_px, expo, conv)
        ow, om = oracle.gpuwarp_eye(np.ascontiguousarray(imgf.transpose(2, 0, 1)), d01, div_px, sep_px, expo, conv)
        if expo in (1.0, 2.0):
            assert np.array_equal(mask[0], om.astype(bool)), (img.shape, div_px, sep_px, expo, conv)
            assert np.abs(warped[0].transpose(2, 0, 1) - ow).max() <= 1e-6
        else:
            assert (mask[0] != om.astype(bool)).mean() <= 1e-3


@pytest.mark.parametrize("sep", [-5.0, 5.0])
def test_large_separation(gu, oracle, sep):
    """Separation at the widget's limits pushes ~40 columns of points off one side of a 795-pixel row: far more segments
    than usual are alive at the first / last output column (the reference's fixed-size active list overflows there --
    DESIGN.md section 7 -- so the oracle, which has no such capacity, is the yardstick)."""
    rng = np.random.default_rng(8)
    h, w = 3, 795
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for kind in ("quant", "noise"):
        d = (syn.make_depth(1, h, w, kind, seed=4, channels=1)[0, ..., 0] * np.float32(255)).astype(np.float32)
        for fill in FILLS:
            got = gu.warp_fill(img, d, fill, -1.0, sep, 2.0, 1.0)[..., :3]
            want = oracle.apply_stereo_divergence(img, d, -1.0, sep, 2.0, fill, 1.0)
            diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
            assert diff.max() <= (1 if fill.startswith('hybrid') else 0), (kind, fill, int((diff > 0).sum()))
