"""-m gpu tests at BASELINE.json's full sizes: the five configs against the oracle where the oracle finishes in
seconds, plus size-independent properties (composition symmetries, frame independence / sharding invariance,
zero-disparity identity) on the big shapes."""
import numpy as np
import pytest
import torch

from comfystereo_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

NODE_DEFAULTS = dict(divergence=3.5, separation=0.0, modes="left-right", stereo_balance=0.0, convergence_point=0.5,
                     stereo_offset_exponent=2.0, depth_blur_edge_threshold=20.0, depth_blur_strength=20.0,
                     depth_map_blur=True, depth_blur_falloff=2.0, depth_blur_vert_smooth=6, batch_size=12)


@pytest.fixture(scope="module")
def node():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from comfystereo_b200 import StereoImageNode
    return StereoImageNode()


def q8(a):
    return np.rint(np.asarray(a) * 255.0).astype(np.uint8)


def run(node, img, dep, **kw):
    p = dict(NODE_DEFAULTS)
    p.update(kw)
    return [o.numpy() for o in node.generate(torch.from_numpy(img), torch.from_numpy(dep), **p)], p


def test_config1_512_naive(node, oracle):
    img, dep = syn.make_image(1, 512, 512, seed=21), syn.make_depth(1, 512, 512, "scene", seed=21)
    got, p = run(node, img, dep, fill_technique="Fill - Naive")
    want = oracle.node_generate(img, dep, **p)
    for g, w_ in zip(got, want):
        assert np.array_equal(q8(g), q8(w_))


@pytest.mark.parametrize("fill", ["Fill - Polylines Sharp", "Fill - Polylines Soft"])
def test_config2_1080p_polylines(node, oracle, fill):
    img, dep = syn.make_image(1, 1080, 1920, seed=22), syn.make_depth(1, 1080, 1920, "scene", seed=22)
    got, p = run(node, img, dep, fill_technique=fill)
    want = oracle.node_generate(img, dep, **p)
    assert got[0].shape == (1, 1080, 3840, 3) and got[3].shape == (1, 1080, 3840)
    for g, w_ in zip(got, want):
        assert np.array_equal(q8(g), q8(w_))


def test_config3_1080p_hybrid_batch(node, oracle):
    img, dep = syn.make_image(3, 1080, 1920, seed=23), syn.make_depth(3, 1080, 1920, "scene", seed=23)
    got, p = run(node, img, dep, fill_technique="Imperfect fill - Hybrid Edge")
    want = oracle.node_generate(img[:1], dep[:1], **p)
    d = np.abs(q8(got[0][:1]).astype(np.int32) - q8(want[0]).astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3          # exp(): CUDA vs libm
    assert np.array_equal(q8(got[1][:1]), q8(want[1])) and np.array_equal(q8(got[2][:1]), q8(want[2]))
    # frames are independent units: the batch equals per-frame runs (what frame sharding relies on)
    one, _ = run(node, img[2:3], dep[2:3], fill_technique="Imperfect fill - Hybrid Edge")
    for a, b in zip(got, one):
        assert np.array_equal(a[2:3], b)


def test_config4_4k_gpuwarp_anaglyph(node, oracle):
    img, dep = syn.make_image(2, 2160, 3840, seed=24), syn.make_depth(2, 2160, 3840, "scene", seed=24)
    got, p = run(node, img, dep, fill_technique="GPU Warp (Fast)", modes="red-cyan-anaglyph", divergence=10.0)
    assert got[0].shape == (2, 2160, 3840, 3) and got[3].shape == (2, 2160, 3840)
    want = oracle.node_generate(img, dep, **p)
    assert np.array_equal(got[3], want[3])                 # unfilled mask: bit-exact
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert np.abs(got[0] - want[0]).max() <= 1e-6


def test_config5_8k_polylines_balance(node, oracle):
    h, w = 3840, 7680
    img, dep = syn.make_image(1, h, w, seed=25), syn.make_depth(1, h, w, "scene", seed=25)
    got, p = run(node, img, dep, fill_technique="Fill - Polylines Soft", stereo_balance=0.5, divergence=4.5)
    assert got[0].shape == (1, h, 2 * w, 3)
    want = oracle.node_generate(img, dep, **p)
    for g, w_ in zip(got, want):
        assert np.array_equal(q8(g), q8(w_))
    # Sharp at this width does not fit the fast kernel's shared-memory tables: the sequential kernel serves it.
    # A band of rows is enough to pin it (rows are independent once the frame's min/max are shared, so the band
    # is cut from the full-frame result of both sides).
    got_s, p = run(node, img, dep, fill_technique="Fill - Polylines Sharp", stereo_balance=0.5, divergence=4.5)
    want_s = oracle.node_generate(img, dep, **p)
    assert np.array_equal(q8(got_s[0]), q8(want_s[0]))


@pytest.mark.parametrize("fill", ["Fill - Polylines Sharp", "GPU Warp (Fast)", "Fill - Naive"])
def test_composition_symmetries_1080p(node, fill):
    img, dep = syn.make_image(2, 1080, 1920, seed=26), syn.make_depth(2, 1080, 1920, "scene", seed=26)
    w, h = 1920, 1080
    lr, _ = run(node, img, dep, fill_technique=fill, modes="left-right")
    rl, _ = run(node, img, dep, fill_technique=fill, modes="right-left")
    tb, _ = run(node, img, dep, fill_technique=fill, modes="top-bottom")
    bt, _ = run(node, img, dep, fill_technique=fill, modes="bottom-top")
    an, _ = run(node, img, dep, fill_technique=fill, modes="red-cyan-anaglyph")
    L, R = lr[0][:, :, :w], lr[0][:, :, w:]
    assert np.array_equal(rl[0][:, :, :w], R) and np.array_equal(rl[0][:, :, w:], L)
    assert np.array_equal(tb[0][:, :h], L) and np.array_equal(tb[0][:, h:], R)
    assert np.array_equal(bt[0][:, :h], R) and np.array_equal(bt[0][:, h:], L)
    assert np.array_equal(an[0][..., 0], L[..., 0]) and np.array_equal(an[0][..., 1:], R[..., 1:])
    for o in (rl, tb, bt, an):   # depth outputs do not depend on the mode
        assert np.array_equal(o[1], lr[1]) and np.array_equal(o[2], lr[2])
    if fill != "GPU Warp (Fast)":   # CPU techniques: mask = pure-black pixels of the composed image
        assert np.array_equal(lr[3], (q8(lr[0]).astype(np.int32).sum(-1) == 0).astype(np.float32))


@pytest.mark.parametrize("fill", ["Fill - Naive", "No fill", "Fill - Naive interpolating"])
def test_flat_depth_is_identity_1080p(node, fill):
    """max == min -> normalised depth 0 -> shift = -conv^exp * div for every pixel; with convergence 0 the shift is
    exactly zero and the forward-warp family must return the truncation-quantised input in both eyes (the
    polylines / splat / reverse-projection techniques resample even at zero disparity)."""
    img = syn.make_image(1, 1080, 1920, seed=27)
    dep = np.full((1, 1080, 1920, 3), 0.37, np.float32)
    got, _ = run(node, img, dep, fill_technique=fill, convergence_point=0.0, depth_map_blur=False)
    want = np.clip(img * np.float32(255), 0, 255).astype(np.uint8)
    assert np.array_equal(q8(got[0][:, :, :1920]), want) and np.array_equal(q8(got[0][:, :, 1920:]), want)


def test_sharding_invariance_gpuwarp_groups(node):
    """GPU Warp couples frames inside a sub-batch (quirk Q9); cutting the batch at sub-batch boundaries, as
    group_aligned_shard_range does, must reproduce the unsharded result exactly."""
    from comfystereo_b200 import engine
    n, bs = 10, 4
    img = syn.make_image(n, 270, 480, seed=28)
    dep = syn.make_depth(n, 270, 480, "scene", seed=28)
    dep[5:] *= np.float32(255.0)   # second half arrives on the 0..255 scale: sub-batch 1 (frames 4-7) is mixed
    full, _ = run(node, img, dep, fill_technique="GPU Warp (Fast)", batch_size=bs, divergence=6.0)
    parts = []
    for r in range(3):
        lo, hi = engine.group_aligned_shard_range(n, r, 3, bs)
        if hi > lo:
            parts.append(run(node, img[lo:hi], dep[lo:hi], fill_technique="GPU Warp (Fast)", batch_size=bs, divergence=6.0)[0])
    for k in range(4):
        assert np.array_equal(full[k], np.concatenate([p_[k] for p_ in parts]))


ALL_FILLS = ['GPU Warp (Fast)', 'No fill', 'No fill - Reverse projection', 'Imperfect fill - Hybrid Edge', 'Fill - Naive',
             'Fill - Naive interpolating', 'Fill - Polylines Soft', 'Fill - Polylines Sharp', 'Fill - Post-fill',
             'Fill - Reverse projection with Post-fill', 'Fill - Hybrid Edge with fill']


@pytest.mark.parametrize("shape", [(37, 53), (5, 7), (1, 33), (9, 2), (64, 1025)])
@pytest.mark.parametrize("fill", ALL_FILLS)
def test_ragged_sizes_against_oracle(node, oracle, shape, fill):
    """Widths that are not multiples of 4 (scalar tails of every vectorised kernel), single rows, two-pixel rows,
    a width just past a power of two -- every technique, every output, against the oracle."""
    h, w = shape
    n = 3
    img = syn.make_image(n, h, w, seed=60 + h + w)
    dep = syn.make_depth(n, h, w, "scene", seed=60 + h + w)
    for mode, blur in (("left-right", True), ("top-bottom", False), ("red-cyan-anaglyph", True)):
        got, p = run(node, img, dep, fill_technique=fill, modes=mode, depth_map_blur=blur, divergence=7.0,
                     separation=0.7, stereo_balance=-0.2, batch_size=2)
        want = oracle.node_generate(img, dep, **p)
        for g, w_ in zip(got, want):
            assert g.shape == w_.shape
        if fill == 'GPU Warp (Fast)':
            assert np.array_equal(got[3], want[3])
            assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
            assert np.abs(got[0] - want[0]).max() <= 1e-6
        else:
            tol = 1 if 'Hybrid' in fill else 0
            assert np.abs(q8(got[0]).astype(np.int32) - q8(want[0]).astype(np.int32)).max() <= tol
            assert np.array_equal(q8(got[1]), q8(want[1])) and np.array_equal(q8(got[2]), q8(want[2]))
            if tol == 0:
                assert np.array_equal(got[3], want[3])


@pytest.mark.parametrize("mode", ["left-right", "right-left", "top-bottom", "bottom-top"])
def test_polylines_composes_in_the_sweep(node, oracle, mode):
    """In the side-by-side / top-bottom modes the Polylines sweep writes the composed float32 tensor and the mask itself.
    Same bytes as the separate k_compose pass (test flag 16), with the sequential kernel (flag 1), tiles (flag 4), and the
    oracle; a black source region makes the mask non-trivial."""
    from comfystereo_b200 import _lib
    img = syn.make_image(2, 96, 352, seed=13, black_box=True)
    dep = syn.make_depth(2, 96, 352, "scene", seed=13)
    lib = _lib.lib()
    for fill in ("Fill - Polylines Sharp", "Fill - Polylines Soft"):
        got, p = run(node, img, dep, fill_technique=fill, modes=mode, divergence=6.0, separation=0.4)
        want = oracle.node_generate(img, dep, **p)
        for g, w_ in zip(got, want):
            assert np.array_equal(q8(g), q8(w_))
        assert np.array_equal(got[3], want[3]) and got[3].sum() > 0
        for flags in (16, 1, 4, 8):
            lib.cs_set_test_flags(flags)
            try:
                alt, _ = run(node, img, dep, fill_technique=fill, modes=mode, divergence=6.0, separation=0.4)
            finally:
                lib.cs_set_test_flags(0)
            for a_, b_ in zip(got, alt):
                assert np.array_equal(a_, b_), flags


@pytest.mark.parametrize("fill", ["No fill", "Fill - Naive", "Fill - Naive interpolating", "No fill - Reverse projection"])
def test_row_techniques_compose_in_the_fill_kernel(node, oracle, fill):
    """The row kernels write the composed tensor and the mask themselves in the side-by-side / top-bottom modes: same
    bytes as the separate k_compose pass (test flag 16) and as the oracle, odd width (no vector path) included."""
    from comfystereo_b200 import _lib
    lib = _lib.lib()
    for w in (352, 333):
        img = syn.make_image(2, 80, w, seed=17, black_box=True)
        dep = syn.make_depth(2, 80, w, "scene", seed=17)
        for mode in ("left-right", "right-left", "top-bottom", "bottom-top"):
            got, p = run(node, img, dep, fill_technique=fill, modes=mode, divergence=7.0, separation=-0.3)
            want = oracle.node_generate(img, dep, **p)
            for g, w_ in zip(got, want):
                assert np.array_equal(q8(g), q8(w_))
            assert np.array_equal(got[3], want[3]) and got[3].sum() > 0
            lib.cs_set_test_flags(16)
            try:
                alt, _ = run(node, img, dep, fill_technique=fill, modes=mode, divergence=7.0, separation=-0.3)
            finally:
                lib.cs_set_test_flags(0)
            for a_, b_ in zip(got, alt):
                assert np.array_equal(a_, b_)


def test_hybrid_edge_composes_in_the_gapfill_kernel(node, oracle):
    """Hybrid Edge writes the composed tensor and the mask from its gap-fill pass when the width is a multiple of 4: same
    bytes as gap fill in place + k_compose (test flag 16; width 333 never fuses), oracle within 1 LSB (exp: DESIGN.md 7)."""
    from comfystereo_b200 import _lib
    lib = _lib.lib()
    for w in (352, 333):
        img = syn.make_image(2, 80, w, seed=19, black_box=True)
        dep = syn.make_depth(2, 80, w, "scene", seed=19)
        for mode in ("left-right", "right-left", "top-bottom", "bottom-top"):
            got, p = run(node, img, dep, fill_technique="Imperfect fill - Hybrid Edge", modes=mode, divergence=7.0, separation=-0.3)
            want = oracle.node_generate(img, dep, **p)
            assert np.abs(q8(got[0]).astype(np.int32) - q8(want[0]).astype(np.int32)).max() <= 1
            assert np.array_equal(q8(got[1]), q8(want[1])) and np.array_equal(q8(got[2]), q8(want[2]))
            assert (got[3] != want[3]).mean() <= 1e-3 and got[3].sum() > 0
            lib.cs_set_test_flags(16)
            try:
                alt, _ = run(node, img, dep, fill_technique="Imperfect fill - Hybrid Edge", modes=mode, divergence=7.0, separation=-0.3)
            finally:
                lib.cs_set_test_flags(0)
            for a_, b_ in zip(got, alt):
                assert np.array_equal(a_, b_)


def test_progress_is_reported_per_chunk_and_depth_batch_may_be_longer(node, oracle, monkeypatch):
    """GS:173 / GS:262: the reference moves its progress bar per sub-batch / per frame; the node reports every chunk of
    frames as its results land (in order, on the calling thread) and the updates add up to the batch size.  A depth
    batch longer than the image batch is simply not read beyond the images (the reference indexes depth_map[i])."""
    import comfystereo_b200.GenerateStereo as gs
    updates = []

    class Bar:
        def __init__(self, total):
            self.total = total

        def update(self, k):
            updates.append(int(k))

    monkeypatch.setattr(gs, "ProgressBar", Bar)
    n = 5
    img = syn.make_image(n, 1080, 1920, seed=3)        # 1080p: the host path moves one frame per chunk
    dep = syn.make_depth(n + 2, 1080, 1920, "scene", seed=3)
    got, p = run(node, img, dep, fill_technique="Fill - Naive", divergence=3.0)
    assert sum(updates) == n and len(updates) >= 2 and all(u > 0 for u in updates), updates
    want = oracle.node_generate(img[:1], dep[:1], **p)
    assert np.array_equal(q8(got[0][:1]), q8(want[0]))
    assert got[0].shape[0] == n


def test_one_process_multi_gpu_sharding(node, oracle, monkeypatch):
    """Inside ComfyUI there is one process: with COMFYSTEREO_MULTI_GPU=1 the node shards the batch frame-wise over every
    visible GPU (one host thread and one pipeline per device, no collective) and assembles in order.  Needs >= 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    monkeypatch.setenv("COMFYSTEREO_MULTI_GPU", "1")
    from comfystereo_b200 import engine
    n = 9
    img = syn.make_image(n, 270, 480, seed=71)
    dep = syn.make_depth(n, 270, 480, "scene", seed=71)
    for fill, bs in (("Fill - Polylines Sharp", 12), ("GPU Warp (Fast)", 2)):
        got, p = run(node, img, dep, fill_technique=fill, batch_size=bs, divergence=6.0)   # multi-GPU path (n >= 2 * ndev)
        key = engine.FILL_NAME_TO_KEY[fill]
        prm = engine.make_params(key, "left-right", 6.0, 0.0, 0.0, 0.5, 2.0, True, 20.0, 20.0, 2.0, 6,
                                 group_size=min(bs, n) if key == "gpu_warp" else 0)
        one = [o.numpy() for o in engine.stereo_batch_host(torch.from_numpy(img), torch.from_numpy(dep), prm, device=0)]
        for a, b in zip(got, one):
            assert np.array_equal(a, b)
    # full-width rows: the Polylines CTAs need the > 48 KB shared-memory opt-in, a per-DEVICE function attribute (a bug once:
    # it was set for the first device only)
    img = syn.make_image(n, 24, 1920, seed=72)
    dep = syn.make_depth(n, 24, 1920, "scene", seed=72)
    for fill in ("Fill - Polylines Sharp", "Fill - Polylines Soft", "Fill - Naive", "Imperfect fill - Hybrid Edge", "GPU Warp (Fast)"):
        got, p = run(node, img, dep, fill_technique=fill, batch_size=3, divergence=4.0)
        key = engine.FILL_NAME_TO_KEY[fill]
        prm = engine.make_params(key, "left-right", 4.0, 0.0, 0.0, 0.5, 2.0, True, 20.0, 20.0, 2.0, 6,
                                 group_size=3 if key == "gpu_warp" else 0)
        one = [o.numpy() for o in engine.stereo_batch_host(torch.from_numpy(img), torch.from_numpy(dep), prm, device=0)]
        for a, b in zip(got, one):
            assert np.array_equal(a, b), fill


def test_4k_polylines_sharp_natural_tiles(node, oracle):
    """3840 px rows do not fit one CTA's tables in sharp mode: the tiled kernel with its natural tile width serves
    them (the 8K config test covers the same path at 7680 px; forced 64-px tiles are covered on every fixture)."""
    img, dep = syn.make_image(1, 2160, 3840, seed=29), syn.make_depth(1, 2160, 3840, "scene", seed=29)
    got, p = run(node, img, dep, fill_technique="Fill - Polylines Sharp", divergence=6.0, separation=-1.0)
    want = oracle.node_generate(img, dep, **p)
    for g, w_ in zip(got, want):
        assert np.array_equal(q8(g), q8(w_))


@pytest.mark.parametrize("fill", ["none", "naive", "naive_interpolating", "inverse", "hybrid_edge", "none_post", "inverse_post"])
def test_very_wide_rows(oracle, fill):
    """Rows of 7700 columns: one row's shared memory leaves room for two CTAs per SM, so the row kernels run with 512-thread
    CTAs there (launch_warp_rows / launch_hybrid); same results as the oracle."""
    import gpu_util as gu
    rng = np.random.default_rng(77)
    for h, w in ((3, 7700), (2, 16384)):     # 16384: a 16K panorama row, close to the 18000-column capacity of these techniques
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        d = (syn.make_depth(1, h, w, "scene", seed=5, channels=1)[0, ..., 0] * np.float32(255)).astype(np.float32)
        got = gu.warp_fill(img, d, fill, 1.5, 0.2, 2.0, 0.5)[..., :3]
        want = oracle.apply_stereo_divergence(img, d, 1.5, 0.2, 2.0, fill, 0.5)
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= (1 if fill == "hybrid_edge" else 0), (fill, w, int((diff > 0).sum()))


@pytest.mark.parametrize("fill,w", [("Fill - Polylines Sharp", 8192), ("Fill - Polylines Soft", 16384),
                                    ("Fill - Polylines Sharp", 12288)])
def test_polylines_wide_rows_are_tiled(oracle, fill, w):
    """The reference has no width limit (SIG:1918-1947 allocates 5 + 2w points per row); Polylines rows of any width are cut
    into tiles here.  8192 px (VR180 8K side by side) and 16384 px rows against the oracle, through the node."""
    from comfystereo_b200 import StereoImageNode
    node = StereoImageNode()
    kw = dict(divergence=3.0, separation=0.5, modes="left-right", stereo_balance=0.3, convergence_point=0.5,
              stereo_offset_exponent=2.0, fill_technique=fill, depth_blur_edge_threshold=20.0, depth_blur_strength=12.0,
              depth_map_blur=True, depth_blur_falloff=2.0, depth_blur_vert_smooth=3, batch_size=2)
    h = 6
    img = syn.make_image(1, h, w, seed=11)
    dep = syn.make_depth(1, h, w, "scene", seed=11, channels=1)
    got = [o.numpy() for o in node.generate(torch.from_numpy(img), torch.from_numpy(dep), **kw)]
    want = oracle.node_generate(img, dep, **kw)
    for g, w_ in zip(got, want):
        assert np.array_equal(q8(g), q8(w_))


def test_polylines_sequential_fallback_with_global_tables(oracle):
    """Rows too wide for the sequential kernel's shared-memory tables (the fallback for rows whose list replay gives up)
    keep them in global scratch: forced here for every row of a 9000 px image, against the oracle."""
    import gpu_util as gu
    rng = np.random.default_rng(5)
    h, w = 3, 9000
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    d = (syn.make_depth(1, h, w, "quant", seed=5, channels=1)[0, ..., 0] * np.float32(255)).astype(np.float32)
    for fill in ("polylines_sharp", "polylines_soft"):
        got = gu.warp_fill(img, d, fill, 2.5, 0.1, 2.0, 0.5, exact=True)[..., :3]
        want = oracle.apply_stereo_divergence(img, d, 2.5, 0.1, 2.0, fill, 0.5)
        assert np.array_equal(got, want), fill


@pytest.mark.parametrize("fill,wmax", [("Fill - Polylines Sharp", 32766), ("GPU Warp (Fast)", 24000), ("Fill - Naive", 18000),
                                       ("Fill - Naive interpolating", 18000), ("Imperfect fill - Hybrid Edge", 18000)])
def test_row_capacity_limits(oracle, fill, wmax):
    """The row techniques keep one row per CTA in shared memory and have a documented maximum width (GPU Warp keeps rows
    beyond ~9200 px in global scratch, up to 24000 px); Polylines is bounded only by the 16-bit point indices of its
    sequential fallback (32766 px sharp, 65533 px soft).  At the limit the node
    still matches the oracle, one step beyond it the library refuses up front (CS_ERR_UNSUPPORTED)."""
    from comfystereo_b200 import StereoImageNode
    from comfystereo_b200._lib import CsError
    node = StereoImageNode()
    kw = dict(divergence=2.0, separation=0.0, modes="left-right", stereo_balance=0.0, convergence_point=0.5,
              stereo_offset_exponent=2.0, fill_technique=fill, depth_blur_edge_threshold=20.0, depth_blur_strength=8.0,
              depth_map_blur=True, depth_blur_falloff=2.0, depth_blur_vert_smooth=2, batch_size=2)
    h = 2
    img = syn.make_image(1, h, wmax, seed=9)
    dep = syn.make_depth(1, h, wmax, "scene", seed=9, channels=1)
    got = [o.numpy() for o in node.generate(torch.from_numpy(img), torch.from_numpy(dep), **kw)]
    want = oracle.node_generate(img, dep, **kw)
    if fill == "GPU Warp (Fast)":
        assert np.array_equal(got[3], want[3]) and np.abs(got[0] - want[0]).max() <= 1e-6
    else:
        q = lambda a: np.rint(a * 255.0).astype(np.int32)
        assert np.abs(q(got[0]) - q(want[0])).max() <= (1 if "Hybrid" in fill else 0)
        assert np.array_equal(got[3], want[3]) or "Hybrid" in fill
    img2 = np.zeros((1, h, wmax + 8, 3), np.float32)
    dep2 = np.zeros((1, h, wmax + 8, 1), np.float32)
    with pytest.raises(CsError, match="exceeds"):
        node.generate(torch.from_numpy(img2), torch.from_numpy(dep2), **kw)


@pytest.mark.parametrize("fill,blur", [("naive", False), ("polylines_sharp", True), ("hybrid_edge", True), ("gpu_warp", True),
                                       ("gpu_warp_mesh", False)])
def test_graph_replay_of_repeated_small_calls(oracle, fill, blur):
    """The second identical small cs_stereo_batch call (same buffers, same parameters) is captured into a CUDA graph and
    later ones replay it: same results as the direct launches, fresh input data is picked up (the graph holds pointers,
    not contents), the launch counter advances by the same amount per call, and a changed parameter is a different graph."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from comfystereo_b200 import engine, _lib
    lib = _lib.lib()
    h, w = 120, 256
    imgs = [torch.from_numpy(syn.make_image(1, h, w, seed=s)).cuda() for s in (1, 2)]
    deps = [torch.from_numpy(syn.make_depth(1, h, w, "scene", seed=s)).cuda() for s in (1, 2)]
    group = 1 if fill.startswith("gpu_warp") else 0
    p = engine.make_params(fill, "left-right", 4.0, 0.5, 0.0, 0.5, 2.0, blur, 9.0, 20.0, 2.0, 2, group_size=group)
    want = []
    for k in range(2):      # fresh tensors every time: different pointers, so these are direct launches
        want.append([o.clone() for o in engine.stereo_batch_device(imgs[k].clone(), deps[k].clone(), p)])
    img, dep = imgs[0].clone(), deps[0].clone()
    out = engine.stereo_batch_device(img, dep, p)
    counts = []
    for rep in range(5):
        k = rep % 2
        img.copy_(imgs[k]); dep.copy_(deps[k])
        lib.cs_launch_count(1)
        engine.stereo_batch_device(img, dep, p, out=out)
        counts.append(lib.cs_launch_count(1))
        torch.cuda.synchronize()
        for a, b in zip(out, want[k]):
            assert torch.equal(a, b), (fill, rep)
    assert len(set(counts)) == 1 and counts[0] > 0, counts
    p2 = engine.make_params(fill, "left-right", 5.0, 0.5, 0.0, 0.5, 2.0, blur, 9.0, 20.0, 2.0, 2, group_size=group)
    for _ in range(3):
        engine.stereo_batch_device(img, dep, p2, out=out)
    torch.cuda.synchronize()
    ref = engine.stereo_batch_device(img.clone(), dep.clone(), p2)
    for a, b in zip(out, ref):
        assert torch.equal(a, b)


@pytest.mark.parametrize("fill,blur", [("polylines_sharp", True), ("naive", False), ("gpu_warp", True)])
def test_hot_path_is_capturable_in_a_callers_cuda_graph(fill, blur):
    """A caller may record cs_stereo_batch into its own CUDA graph (torch.cuda.graph): every launch of the chunk sequence is
    stream-ordered with no host synchronisation, and the library's own graph cache steps aside while the stream captures."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from comfystereo_b200 import engine
    h, w = 96, 200
    img = torch.from_numpy(syn.make_image(2, h, w, seed=5)).cuda()
    dep = torch.from_numpy(syn.make_depth(2, h, w, "scene", seed=5)).cuda()
    p = engine.make_params(fill, "left-right", 5.0, 0.3, 0.0, 0.5, 2.0, blur, 9.0, 20.0, 2.0, 2,
                           group_size=2 if fill == "gpu_warp" else 0)
    want = [o.clone() for o in engine.stereo_batch_device(img, dep, p)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):                      # includes the calls that would build the library's own graph
            engine.stereo_batch_device(img, dep, p)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = engine.stereo_batch_device(img, dep, p)
    for o in outs:
        o.zero_()
    g.replay()
    torch.cuda.synchronize()
    for a, b in zip(outs, want):
        assert torch.equal(a, b)
    img2 = torch.from_numpy(syn.make_image(2, h, w, seed=6)).cuda()
    want2 = [o.clone() for o in engine.stereo_batch_device(img2, dep, p)]
    img.copy_(img2)
    g.replay()
    torch.cuda.synchronize()
    for a, b in zip(outs, want2):
        assert torch.equal(a, b)


def test_graph_cache_eviction_and_release():
    """More distinct small jobs than the graph cache holds (16): entries are captured, evicted and destroyed while results
    stay those of direct launches; cs_host_release drops whatever is cached and the next calls start over."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from comfystereo_b200 import engine
    p = engine.make_params("naive", "left-right", 4.0, 0.0, 0.0, 0.5, 2.0, False, 9.0, 20.0, 2.0, 2)
    jobs = []
    for k in range(20):
        h, w = 16 + 2 * k, 64 + 8 * k
        img = torch.from_numpy(syn.make_image(1, h, w, seed=k)).cuda()
        dep = torch.from_numpy(syn.make_depth(1, h, w, "scene", seed=k)).cuda()
        want = [o.clone() for o in engine.stereo_batch_device(img.clone(), dep.clone(), p)]
        out = engine.stereo_batch_device(img, dep, p)
        jobs.append((img, dep, out, want))
    for rounds in range(3):
        for img, dep, out, want in jobs:
            engine.stereo_batch_device(img, dep, p, out=out)
        torch.cuda.synchronize()
        for img, dep, out, want in jobs:
            for a, b in zip(out, want):
                assert torch.equal(a, b)
        if rounds == 1:
            engine.release()
