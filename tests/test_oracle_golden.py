"""Pins the CPU oracle (oracle/stereo_oracle.c + oracle/oracle.py) against fixtures produced by
the UNMODIFIED reference (oracle/make_golden.py).  CPU only.

Tolerances (stated once, used everywhere):
  * integer work (uint8 images given identical depth, source-column indices, masks): bit-exact
  * blur (float32, torch conv2d summation order is unspecified): <= 2e-4 on the 0..255 scale
  * GPU-Warp float image: <= 2e-5 (torch.linspace rounding gives rows a ~1e-6 blend)
  * node level: with the reference's captured blurred depth injected (stage-wise protocol,
    SURVEY.md section 8) every output is bit-exact; the oracle's own blur is held to 1 LSB on
    the (wrapping, quirk Q1) depth outputs.
"""
import json
import os
import sys
import zlib

import numpy as np
import pytest

from conftest import load_manifest, load_golden, circ_dist_u8
from comfystereo_b200 import synthetic as syn

MAN = load_manifest()


def _crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def _node_inputs(spec):
    img = syn.make_image(spec["n"], spec["h"], spec["w"], seed=spec["seed"], black_box=spec["black_box"])
    dep = syn.make_depth(spec["n"], spec.get("dh", spec["h"]), spec.get("dw", spec["w"]), spec["kind"],
                         seed=spec["seed"], channels=spec["channels"], scale255=spec["scale255"])
    return img, dep


@pytest.mark.parametrize("spec", MAN["stage"], ids=[s["name"] + "_" + s["kind"] for s in MAN["stage"]])
def test_stage_vs_reference(oracle, spec):
    g = load_golden("stage", spec["name"])
    h, w = spec["h"], spec["w"]
    d = syn.make_depth(1, h, w, spec["kind"], seed=spec["seed"])[0, ..., 0]
    if spec["stage"] == "blur":
        d255 = d * np.float32(255)
        assert _crc(d255) == int(g["crc"])
        L, R = oracle.blur(d255, spec["strength"], spec["thr"], spec["falloff"], spec["vert"])
        assert np.abs(L - g["L"]).max() <= 2e-4
        assert np.abs(R - g["R"]).max() <= 2e-4
    elif spec["stage"] == "warp":
        probe = syn.index_probe_image(h, w)
        d255 = d * np.float32(255)
        assert _crc(probe, d255) == int(g["crc"])
        out = oracle.apply_stereo_divergence(probe, d255, spec["div"], spec["sep"], spec["expo"],
                                             spec["fill"], spec["conv"])
        assert np.array_equal(out, g["out"])  # bit-exact, includes the decoded source columns
    else:
        img = syn.make_image(1, h, w, seed=spec["seed"])
        assert _crc(img, d) == int(g["crc"])
        chw = np.ascontiguousarray(img[0].transpose(2, 0, 1))
        warped, mask = oracle.gpuwarp_eye(chw, d, spec["div_px"], spec["sep_px"], spec["expo"], spec["conv"])
        assert np.array_equal(mask.astype(np.uint8), g["mask"])
        assert np.abs(warped - g["warped"]).max() <= 2e-5


@pytest.mark.parametrize("spec", MAN["dark"], ids=[s["name"] + "_" + s["fill"] for s in MAN["dark"]])
def test_dark_images_vs_reference(oracle, spec):
    """apply_stereo_divergence on dark images (black ramp values, black-but-filled pixels; the interpolating fill's
    float32 ramp arithmetic): bit-exact against the reference."""
    g = load_golden("dark", spec["name"])
    img, d = syn.dark_case(spec["seed"])
    assert _crc(img, d) == int(g["crc"])
    out = oracle.apply_stereo_divergence(img, d, spec["div"], spec["sep"], 1.0, spec["fill"], 0.5)
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("spec", MAN["resize"], ids=[s["name"] for s in MAN["resize"]])
def test_resize_vs_reference(oracle, spec):
    """N1 depth resize (GS:141-148 / GS:214-220): bit-exact against torch's strict (ATEN_CPU_CAPABILITY=default)
    kernel; torch's FMA-contracting AVX builds stay within ~1 ulp of the source coordinate of it."""
    g = load_golden("resize", spec["name"])
    d = syn.make_depth(1, spec["dh"], spec["dw"], spec["kind"], seed=spec["seed"], channels=1)[0, ..., 0]
    assert _crc(d) == int(g["crc"])
    out = oracle.resize_bilinear(d, (spec["h"], spec["w"]))
    assert np.array_equal(out, g["out_strict"])
    assert np.abs(out - g["out_native"]).max() <= 1e-4


@pytest.mark.parametrize("spec", MAN["node"], ids=[s["name"] for s in MAN["node"]])
def test_node_vs_reference(oracle, spec):
    """End to end through the node glue.  When the blur is on, the reference's own blurred
    depth (captured in the fixture) is injected so that everything downstream is held to the
    integer-exact bar; the blur itself is pinned by the stage tests above and re-checked here."""
    g = load_golden("node", spec["name"])
    img, dep = _node_inputs(spec)
    assert _crc(img, dep) == int(g["crc"])
    override = (g["blur_l"], g["blur_r"]) if "blur_l" in g.files else None
    stereo, dl, dr, mask = oracle.node_generate(img, dep, blur_override=override, **spec["params"])
    is_gw = spec["params"]["fill_technique"] == 'GPU Warp (Fast)'
    if is_gw:
        assert stereo.shape == g["stereo"].shape
        assert np.array_equal((mask > 0).astype(np.uint8), g["mask"])
        assert np.array_equal(dl[..., 0], g["depth_l"]) and np.array_equal(dr[..., 0], g["depth_r"])
        assert np.abs(stereo - g["stereo"]).max() <= 2e-5
    else:
        q = lambda a: np.rint(a * 255.0).astype(np.uint8)
        assert np.array_equal(q(stereo), g["stereo"])
        assert np.array_equal(q(dl[..., 0]), g["depth_l"]) and np.array_equal(q(dr[..., 0]), g["depth_r"])
        assert np.array_equal(q(mask), g["mask"])
    if override is not None:  # the oracle's own blur, same inputs, float tolerance
        own = oracle.node_generate(img, dep, **spec["params"])
        # ... and end to end: exactly the measured number of pixels the blur's rounding noise moves (0 on continuous depth)
        sys_path_oracle = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
        if sys_path_oracle not in sys.path:
            sys.path.insert(0, sys_path_oracle)
        from measure_node_flips import flips as count_flips
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "node_flip_counts.json")) as f:
            rec = json.load(f)[spec["name"]]
        px, mk = count_flips(spec, g, own[0], own[3])
        assert (px, mk) == (rec["pixels"], rec["mask"]), (px, mk, rec)
        if spec["kind"] in ("scene", "noise"):
            assert px == 0 and mk == 0
        if is_gw:
            assert np.abs(own[1][..., 0] - g["depth_l"]).max() <= 1e-6
            assert np.abs(own[2][..., 0] - g["depth_r"]).max() <= 1e-6
        else:
            assert circ_dist_u8(q(own[1][..., 0]), g["depth_l"]).max() <= 1
            assert circ_dist_u8(q(own[2][..., 0]), g["depth_r"]).max() <= 1


def _array_inputs(spec):
    img = (syn.make_image(1, spec["h"], spec["w"], seed=spec["seed"])[0] * 255).astype(np.uint8)
    d = (syn.make_depth(1, spec["h"], spec["w"], spec["kind"], seed=spec["seed"], channels=1)[0, ..., 0] * np.float32(255))
    return img, d.astype(np.float32)


@pytest.mark.parametrize("spec", MAN.get("arrays", []), ids=[s["name"] + "_" + s["fill"] for s in MAN.get("arrays", [])])
def test_array_inputs_vs_reference(oracle, spec):
    """create_stereoimages with NON-tensor inputs (numpy arrays / PIL, SIG:1486-1496): the scipy blur (reflected Sobel,
    nearest-border float64 box sums in scipy's own order) is restated bit for bit, so every output is exact."""
    g = load_golden("arrays", spec["name"])
    img, d = _array_inputs(spec)
    assert _crc(img, d) == int(g["crc"])
    if spec["blur"]:
        bl, br = oracle.blur_numpy(d, spec["s"], spec["thr"], spec["fo"], spec["v"])
        assert np.array_equal(bl, g["blur_l"], equal_nan=True) and np.array_equal(br, g["blur_r"], equal_nan=True)
    res, dl, dr = oracle.create_stereoimages_arrays(img, d, spec["div"], spec["sep"], list(spec["modes"]), spec["bal"], spec["expo"],
                                                    spec["fill"], spec["s"], spec["thr"], spec["blur"], spec["conv"], spec["fo"],
                                                    spec["v"])
    for i, r in enumerate(res):
        diff = np.abs(r.astype(np.int32) - g[f"stereo{i}"].astype(np.int32))
        assert diff.max() <= (1 if spec["fill"].startswith("hybrid") else 0)
    assert np.array_equal(dl, g["depth_l"])
    if spec["blur"]:
        assert np.array_equal(dr, g["depth_r"])
