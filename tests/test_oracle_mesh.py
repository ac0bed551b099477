"""CPU: the mesh-warp rule set of the oracle (oracle/stereo_oracle.c:orc_mesh_offsets / orc_mesh_raster, the restatement of
forward_warp_mesh, SIG:453-689).

The reference draws this mesh with OpenGL; fragment coverage and interpolation rounding are the GL implementation's, so
there are no golden vectors to pin against ("parity unpinned").  What CAN be checked without a GL
context is that the rule set is a faithful rasterisation of the reference's mesh: an independent float64 rasteriser written
from the OpenGL rules alone (pixel-centre sampling of each kept triangle with edge functions, barycentric interpolation,
nearest depth wins) must agree with it everywhere except on pixels that sit on a triangle edge or a depth tie to within
float32 rounding."""
import numpy as np
import pytest

from comfystereo_b200 import synthetic as syn


def _offsets64(d01, div_px, sep_px, expo, conv):
    d = d01.astype(np.float64)
    rng = d.max() - d.min()
    nd = (d - d.min()) / max(rng, 1e-6) if rng > 1e-6 else np.zeros_like(d)
    sh = nd - conv
    return nd, np.sign(sh) * np.abs(sh) ** expo * div_px + sep_px


def _gl_raster64(img_chw, d01_b, b, div_px, sep_px, expo, conv):
    """Frame b of the sub-batch d01_b, the way the OpenGL pipeline is specified: window coords from clip coords, a pixel is
    covered by a triangle when its centre is inside (edge functions), attributes are barycentric, smaller clip_z wins."""
    B, H, W = d01_b.shape
    nds, pos = zip(*[_offsets64(d01_b[k], div_px, sep_px, expo, conv) for k in range(B)])
    tris = []
    for r in range(H - 1):
        for c in range(W - 1):
            v00, v10, v01, v11 = (r, c), (r, c + 1), (r + 1, c), (r + 1, c + 1)
            tris.append((v00, v10, v01))
            tris.append((v11, v10, v01))
    keep = []
    for t in tris:
        ok = False
        for k in range(B):
            o = [pos[k][v] for v in t]
            ok |= max(abs(o[0] - o[1]), abs(o[0] - o[2]), abs(o[1] - o[2])) < 1.5
        keep.append(ok)
    nd, po = nds[b], pos[b]
    cols, rows = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    clip_x = (cols + po) / (W - 1) * 2 - 1
    clip_y = -(rows / (H - 1) * 2 - 1)
    X = (clip_x + 1) * W / 2
    Y = H - (clip_y + 1) * H / 2          # row 0 of the flipped read-back is the top of the GL window
    out = np.zeros((3, H, W))
    zb = np.full((H, W), -np.inf)
    cov = np.zeros((H, W), bool)
    margin = np.full((H, W), np.inf)       # how close the decision at this pixel was (edge distance / z gap)
    px, py = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    for t, k in zip(tris, keep):
        if not k:
            continue
        (r0, c0), (r1, c1), (r2, c2) = t
        x0, y0, x1, y1, x2, y2 = X[r0, c0], Y[r0, c0], X[r1, c1], Y[r1, c1], X[r2, c2], Y[r2, c2]
        area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        if area == 0:
            continue
        jlo, jhi = max(int(np.floor(min(y0, y1, y2))) - 1, 0), min(int(np.ceil(max(y0, y1, y2))) + 1, H)
        ilo, ihi = max(int(np.floor(min(x0, x1, x2))) - 1, 0), min(int(np.ceil(max(x0, x1, x2))) + 1, W)
        if jlo >= jhi or ilo >= ihi:
            continue
        sx, sy = px[jlo:jhi, ilo:ihi], py[jlo:jhi, ilo:ihi]
        w0 = ((x1 - sx) * (y2 - sy) - (x2 - sx) * (y1 - sy)) / area
        w1 = ((x2 - sx) * (y0 - sy) - (x0 - sx) * (y2 - sy)) / area
        w2 = 1 - w0 - w1
        inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
        edge = np.minimum(np.minimum(np.abs(w0), np.abs(w1)), np.abs(w2))
        z = w0 * nd[r0, c0] + w1 * nd[r1, c1] + w2 * nd[r2, c2]
        sub_m = margin[jlo:jhi, ilo:ihi]
        near = edge < 1e-4                  # the centre is (almost) on an edge: either answer is a valid rasterisation
        sub_m[near] = 0
        sub_z, sub_c = zb[jlo:jhi, ilo:ihi], cov[jlo:jhi, ilo:ihi]
        gap = np.abs(z - sub_z)
        sub_m[inside & sub_c] = np.minimum(sub_m[inside & sub_c], gap[inside & sub_c])
        win = inside & (z > sub_z)
        for ch in range(3):
            col = w0 * img_chw[ch, r0, c0] + w1 * img_chw[ch, r1, c1] + w2 * img_chw[ch, r2, c2]
            out[ch, jlo:jhi, ilo:ihi][win] = col[win]
        sub_z[win] = z[win]
        sub_c |= inside
    return out, ~cov, margin


@pytest.mark.parametrize("seed,h,w,bsz,div,expo", [(1, 12, 17, 1, 3.0, 1.0), (2, 9, 24, 2, -4.0, 2.0), (3, 16, 16, 3, 6.0, 0.5),
                                                 (4, 7, 31, 1, -2.5, 3.0)])
def test_rule_set_is_a_rasterisation_of_the_mesh(oracle, seed, h, w, bsz, div, expo):
    rng = np.random.default_rng(seed)
    img = rng.random((bsz, 3, h, w), dtype=np.float32)
    d = np.stack([syn.make_depth(1, h, w, seed=seed * 10 + k)[0, :, :, 0] for k in range(bsz)]).astype(np.float32)
    d = d / np.float32(max(d.max(), 1.0))
    div_px, sep_px, conv = div / 100 * w * 4, 0.3, 0.5      # a few pixels of shift at this tiny width
    out, mask = oracle.meshwarp_batch(img, d, div_px, sep_px, expo, conv)
    total = decided = 0
    for b in range(bsz):
        ref, rmask, margin = _gl_raster64(img[b].astype(np.float64), d, b, div_px, sep_px, expo, conv)
        clear = margin > 1e-4                      # pixels whose outcome no rounding can change
        total += clear.size
        decided += int(clear.sum())
        assert np.array_equal(mask[b][clear], rmask[clear])
        cov = clear & ~rmask
        assert np.abs(out[b][:, cov] - ref[:, cov]).max() <= 2e-5
    assert decided >= 0.8 * total                  # the comparison is not vacuous


def test_mesh_properties(oracle):
    rng = np.random.default_rng(5)
    h, w = 20, 33
    d = rng.random((2, h, w), dtype=np.float32)
    # a constant image comes out constant wherever anything is drawn or smeared, 0 elsewhere
    img = np.full((2, 3, h, w), 0.375, np.float32)
    out, mask = oracle.meshwarp_batch(img, d, 5.0, 0.0, 1.0, 0.5)
    assert set(np.unique(out)) <= {np.float32(0.0), np.float32(0.375)}
    assert (out[:, 0][~mask] == np.float32(0.375)).all()
    # flat depth, no separation: one rigid shift of the whole sheet, every pixel it still reaches is covered
    flat = np.full((1, h, w), 0.5, np.float32)
    img = rng.random((1, 3, h, w), dtype=np.float32)
    out, mask = oracle.meshwarp_batch(img, flat, 7.0, 0.0, 1.0, 0.0)     # nd = 0 everywhere, conv 0 -> offset 0
    assert not mask.any()
    assert out.min() >= img.min() - 1e-6 and out.max() <= img.max() + 1e-6
    # values stay inside the hull of the image whatever the depth
    out, mask = oracle.meshwarp_batch(img, d[:1], -6.0, 1.0, 2.0, 0.3)
    assert out.min() >= 0.0 and out.max() <= img.max() + 1e-6
    # left eye (divergence >= 0) smears from the left: a gap pixel repeats its left neighbour
    out, mask = oracle.meshwarp_batch(img, d[:1], 9.0, 0.0, 1.0, 0.5)
    j, i = np.nonzero(mask[0][:, 1:])
    assert len(j) and np.array_equal(out[0][:, j, i + 1], out[0][:, j, i])
    out, mask = oracle.meshwarp_batch(img, d[:1], -9.0, 0.0, 1.0, 0.5)
    j, i = np.nonzero(mask[0][:, :-1])
    assert len(j) and np.array_equal(out[0][:, j, i], out[0][:, j, i + 1])


def test_mesh_topology_is_sub_batch_wide(oracle):
    """SIG:536: a triangle is kept when it passes the gradient test in ANY frame, so a noisy frame drawn next to a flat one
    keeps every triangle -- and has no gaps other than the sheet's border."""
    rng = np.random.default_rng(9)
    h, w = 16, 40
    noisy = rng.random((h, w), dtype=np.float32)
    img = rng.random((2, 3, h, w), dtype=np.float32)
    _, alone = oracle.meshwarp_batch(img[:1], noisy[None], 8.0, 0.0, 1.0, 0.5)
    _, paired = oracle.meshwarp_batch(img, np.stack([noisy, np.full((h, w), 0.5, np.float32)]), 8.0, 0.0, 1.0, 0.5)
    assert alone[0].sum() > paired[0].sum()
    assert alone[0][:, 5:-5].any() and not paired[0][:, 5:-5].any()


def test_mesh_degenerate_sizes(oracle):
    img = np.ones((1, 3, 1, 8), np.float32)
    out, mask = oracle.meshwarp_batch(img, np.zeros((1, 1, 8), np.float32), 3.0, 0.0, 1.0, 0.5)
    assert mask.all() and not out.any()          # a single row has no triangles: SIG:506-520 builds (H-1)*(W-1) quads
    img = np.ones((1, 3, 8, 1), np.float32)
    out, mask = oracle.meshwarp_batch(img, np.zeros((1, 8, 1), np.float32), 3.0, 0.0, 1.0, 0.5)
    assert mask.all() and not out.any()


def test_pipeline_with_mesh_warp(oracle):
    """create_stereoimages_gpu(mesh=True): composition, masks and depth outputs are the scatter path's; only warp_fn changes."""
    h, w, b = 24, 48, 2
    img = syn.make_image(b, h, w, seed=3).transpose(0, 3, 1, 2).copy()
    d = syn.make_depth(b, h, w, seed=4)[..., 0] / np.float32(255)
    for mode in ['left-right', 'top-bottom', 'red-cyan-anaglyph', 'only-right']:
        res, dl, dr, mask = oracle.create_stereoimages_gpu(img, d, 4.0, 1.0, [mode], stereo_balance=0.25, mesh=True)
        res0, dl0, dr0, _ = oracle.create_stereoimages_gpu(img, d, 4.0, 1.0, [mode], stereo_balance=0.25)
        assert res[0].shape == res0[0].shape and mask.shape == (b, h, w)
        assert np.array_equal(dl, dl0) and np.array_equal(dr, dr0)
        assert np.isfinite(res[0]).all()
