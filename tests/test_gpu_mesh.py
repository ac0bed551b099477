"""-m gpu: the mesh warp (forward_warp_mesh, SIG:453-689; CS_FILL_GPU_WARP_MESH / cs_forward_warp_mesh) against the oracle's
rule set (oracle/stereo_oracle.c:orc_mesh_raster, itself checked against an independent OpenGL-rule rasteriser in
tests/test_oracle_mesh.py).  Parity with the REFERENCE is unpinned for this path -- its OpenGL rasterisation is
implementation-defined -- so the bar here is CUDA == oracle, bit for bit: same float32 operations in the same order."""
import numpy as np
import pytest
import torch

from comfystereo_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

SPECIAL = (1.0, 2.0, 0.5, 3.0)      # exponents torch.pow / the kernels / the oracle evaluate without powf


@pytest.fixture(scope="module")
def gu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gpu_util
    return gpu_util


def _warp(img_bhwc, d, div_px, sep_px, expo, conv):
    from comfystereo_b200 import engine
    w, m = engine.forward_warp_device(torch.from_numpy(np.ascontiguousarray(img_bhwc)).cuda(),
                                      torch.from_numpy(np.ascontiguousarray(d)).cuda(), div_px, sep_px, expo, conv, mesh=True)
    torch.cuda.synchronize()
    return w.cpu().numpy(), m.cpu().numpy() > 0.5


@pytest.mark.parametrize("h,w,b,kind,div_px,sep_px,expo,conv", [
    (32, 64, 1, "scene", 6.0, 0.0, 1.0, 0.5),
    (33, 127, 3, "scene", -9.5, 1.25, 2.0, 0.5),
    (40, 200, 2, "noise", 12.0, -2.0, 0.5, 0.3),
    (17, 90, 4, "quant", -7.0, 0.0, 3.0, 0.7),
    (2, 2, 1, "noise", 1.0, 0.0, 1.0, 0.5),
    (2, 300, 2, "scene", 25.0, 0.0, 1.0, 0.0),
    (64, 3, 1, "noise", 2.0, 0.5, 2.0, 1.0),
    (270, 480, 2, "scene", 21.6, 0.0, 2.0, 0.5),
    (25, 60, 2, "flat", 5.0, 1.0, 1.0, 0.5),
])
def test_forward_warp_mesh_bit_exact(gu, oracle, h, w, b, kind, div_px, sep_px, expo, conv):
    img = syn.make_image(b, h, w, seed=h + w)
    d = syn.make_depth(b, h, w, kind, seed=h * 3 + w)[..., 0]
    d = (d / np.float32(255)).astype(np.float32) if d.max() > 1 else d
    warped, mask = _warp(img, d, div_px, sep_px, expo, conv)
    ow, om = oracle.meshwarp_batch(img.transpose(0, 3, 1, 2), d, div_px, sep_px, expo, conv)
    assert np.array_equal(mask, om)
    assert np.array_equal(warped.transpose(0, 3, 1, 2), ow)


def test_forward_warp_mesh_scale_rule_and_degenerate(gu, oracle):
    """SIG:487-489: /255 when ANY frame's max > 1 (whole batch).  A 1-row frame has no triangles at all."""
    img = syn.make_image(2, 20, 50, seed=1)
    d = syn.make_depth(2, 20, 50, "scene", seed=2)[..., 0].astype(np.float32)
    d = d / np.float32(max(d.max(), 1.0))
    d[1] *= np.float32(255)                              # one frame on the 0..255 scale -> both are divided
    warped, mask = _warp(img, d, 8.0, 0.0, 1.0, 0.5)
    ow, om = oracle.meshwarp_batch(img.transpose(0, 3, 1, 2), d / np.float32(255), 8.0, 0.0, 1.0, 0.5)
    assert np.array_equal(mask, om) and np.array_equal(warped.transpose(0, 3, 1, 2), ow)
    img1 = syn.make_image(1, 1, 40, seed=3)
    warped, mask = _warp(img1, np.zeros((1, 1, 40), np.float32), 3.0, 0.0, 1.0, 0.5)
    assert mask.all() and not warped.any()


@pytest.mark.parametrize("seed", range(4))
def test_random_mesh_cases(gu, oracle, seed):
    rng = np.random.default_rng(3000 + seed)
    for _ in range(10):
        img, d, div, sep, expo, conv = syn.fuzz_case(rng)
        h, w = d.shape
        b = int(rng.integers(1, 4))
        imgf = np.stack([np.roll(img, k, axis=1) for k in range(b)]).astype(np.float32) / np.float32(255)
        d01 = np.stack([np.roll(d, 2 * k, axis=0) for k in range(b)]).astype(np.float32) / np.float32(255)
        div_px, sep_px = div / 100.0 * w, sep / 100.0 * w
        warped, mask = _warp(imgf, d01, div_px, sep_px, expo, conv)
        ow, om = oracle.meshwarp_batch(imgf.transpose(0, 3, 1, 2), d01, div_px, sep_px, expo, conv)
        if expo in SPECIAL:
            assert np.array_equal(mask, om), (imgf.shape, div_px, sep_px, expo, conv)
            assert np.array_equal(warped.transpose(0, 3, 1, 2), ow), (imgf.shape, div_px, sep_px, expo, conv)
        else:   # powf: CUDA vs libm differ by an ulp, which can move a pixel centre across an edge
            assert (mask != om).mean() <= 5e-3, (imgf.shape, div_px, sep_px, expo, conv)


MODES = ["left-right", "right-left", "top-bottom", "bottom-top", "red-cyan-anaglyph", "left-only", "only-right",
         "cyan-red-reverseanaglyph"]


@pytest.mark.parametrize("blur", [False, True])
def test_create_stereoimages_gpu_with_mesh_warp(gu, oracle, monkeypatch, blur):
    """The function-level pipeline with the reference's module switch on: every mode, unbalanced eyes, sub-batch of 3."""
    from comfystereo_b200 import stereoimage_generation as sig
    monkeypatch.setattr(sig, "MODERNGL_AVAILABLE", True)
    b, h, w = 3, 54, 96
    img = syn.make_image(b, h, w, seed=11).transpose(0, 3, 1, 2).copy()
    d = (syn.make_depth(b, h, w, "scene", seed=12)[..., 0] / np.float32(255)).astype(np.float32)
    kw = dict(stereo_balance=0.3, stereo_offset_exponent=2.0, convergence_point=0.4, depth_blur_strength=7.0,
              depth_blur_edge_threshold=10.0, direction_aware_depth_blur=blur, depth_blur_falloff=2.0, depth_blur_vert_smooth=2)
    res, dl, dr, mask = sig.create_stereoimages_gpu(torch.from_numpy(img), torch.from_numpy(d), 5.0, 0.8, MODES, **kw)
    ores, odl, odr, omask = oracle.create_stereoimages_gpu(img, d, 5.0, 0.8, MODES, mesh=True, **kw)
    assert np.array_equal(dl.cpu().numpy(), odl) and np.array_equal(dr.cpu().numpy(), odr)
    assert np.array_equal(mask.cpu().numpy(), omask)
    for mode, a, o in zip(MODES, res, ores):
        assert np.array_equal(a.cpu().numpy(), o), mode
    # passthrough eye (stereo_balance 1.0: right divergence 0) and the scatter warp untouched by the switch being off
    res, _, _, mask = sig.create_stereoimages_gpu(torch.from_numpy(img), torch.from_numpy(d), 5.0, 0.0, ["left-right"],
                                                 stereo_balance=1.0)
    ores, _, _, omask = oracle.create_stereoimages_gpu(img, d, 5.0, 0.0, ["left-right"], stereo_balance=1.0, mesh=True)
    assert np.array_equal(res[0].cpu().numpy(), ores[0]) and np.array_equal(mask.cpu().numpy(), omask)
    monkeypatch.setattr(sig, "MODERNGL_AVAILABLE", False)
    res, _, _, _ = sig.create_stereoimages_gpu(torch.from_numpy(img), torch.from_numpy(d), 5.0, 0.0, ["left-right"])
    ores, _, _, _ = oracle.create_stereoimages_gpu(img, d, 5.0, 0.0, ["left-right"])
    assert np.abs(res[0].cpu().numpy() - ores[0]).max() <= 1e-6


def test_node_with_mesh_warp(gu, oracle, monkeypatch):
    """Through the node (host tensors -> cs_stereo_batch_host), sub-batches of 4 over 10 frames: the topology is per sub-batch."""
    from comfystereo_b200 import stereoimage_generation as sig
    from comfystereo_b200.GenerateStereo import StereoImageNode
    monkeypatch.setattr(sig, "MODERNGL_AVAILABLE", True)
    n, h, w = 10, 45, 80
    img = syn.make_image(n, h, w, seed=31)
    dep = syn.make_depth(n, h, w, "scene", seed=32)
    dep = (dep / np.float32(255)).astype(np.float32) if dep.max() > 1 else dep
    params = dict(divergence=6.0, separation=0.5, modes="left-right", stereo_balance=0.0, convergence_point=0.5,
                  stereo_offset_exponent=2.0, fill_technique="GPU Warp (Fast)", depth_blur_edge_threshold=20.0,
                  depth_blur_strength=6.0, depth_map_blur=True, depth_blur_falloff=1.0, depth_blur_vert_smooth=1, batch_size=4)
    got = [o.numpy() for o in StereoImageNode().generate(torch.from_numpy(img), torch.from_numpy(dep), **params)]
    want = oracle.node_generate(img, dep, mesh=True, **params)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert np.array_equal(got[3], want[3])
    assert np.array_equal(got[0], want[0])
    # and it is not the scatter warp
    monkeypatch.setattr(sig, "MODERNGL_AVAILABLE", False)
    other = StereoImageNode().generate(torch.from_numpy(img), torch.from_numpy(dep), **params)[0].numpy()
    assert not np.array_equal(other, got[0])


def test_mesh_width_limit_and_wide_rows(gu, oracle):
    from comfystereo_b200 import engine, _lib
    # 8192: > 48 KB of shared memory per row (the opt-in path, 512-thread CTAs); 16384: the row state no longer fits a CTA's
    # shared memory and lives in global scratch (k_meshwarp_wide) -- a 16K panorama row
    for h, w, div_px in ((4, 8192, 160.0), (3, 16384, 300.0)):
        img = syn.make_image(2, h, w, seed=5)
        d = (syn.make_depth(2, h, w, "scene", seed=6)[..., 0] / np.float32(255)).astype(np.float32)
        warped, mask = _warp(img, d, div_px, 0.0, 1.0, 0.5)
        ow, om = oracle.meshwarp_batch(img.transpose(0, 3, 1, 2), d, div_px, 0.0, 1.0, 0.5)
        assert np.array_equal(mask, om) and np.array_equal(warped.transpose(0, 3, 1, 2), ow), w
    with pytest.raises(_lib.CsError, match="24000"):
        engine.forward_warp_device(torch.zeros(1, 2, 24001, 3).cuda(), torch.zeros(1, 2, 24001).cuda(), 1.0, 0.0, 1.0, 0.5, mesh=True)


def test_scatter_warp_16k_rows_in_global_scratch(gu, oracle):
    """forward_warp_gpu on rows wider than a CTA's shared memory can hold (k_gpuwarp_wide): same result as the oracle."""
    h, w = 3, 16384
    img = syn.make_image(2, h, w, seed=8)
    d = (syn.make_depth(2, h, w, "scene", seed=9)[..., 0] / np.float32(255)).astype(np.float32)
    warped, mask = gu.forward_warp(img, d, 250.0, -3.0, 2.0, 0.5)
    for b in range(2):
        ow, om = oracle.gpuwarp_eye(np.ascontiguousarray(img[b].transpose(2, 0, 1)), d[b], 250.0, -3.0, 2.0, 0.5)
        assert np.array_equal(mask[b], om)
        assert np.abs(warped[b].transpose(2, 0, 1) - ow).max() <= 1e-6
