"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not exist on the
GPU box):      python oracle/make_golden.py

Two fixture families, both small enough for git (inputs are regenerated from seeds by
comfystereo_b200/synthetic.py and guarded by a CRC stored in the fixture):

  node_*.npz   StereoImageNode.generate(...) end to end (GS:79-353): the 4 returned tensors.
               CPU techniques are stored as uint8 (the node returns uint8/255 exactly),
               'GPU Warp (Fast)' as float32.
  stage_*.npz  single hot-path functions called directly on explicit inputs:
               directional_motion_blur_gpu (SIG:1171), apply_stereo_divergence_* (SIG:1715-1992)
               on an index-probe image (exact source-column view), forward_warp_gpu (SIG:277).
"""
import json
import os
import sys
import warnings
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from comfystereo_b200 import synthetic as syn  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

DEFAULTS = dict(divergence=4.5, separation=0.0, modes="left-right", stereo_balance=0.0,
                convergence_point=0.5, stereo_offset_exponent=2.0, fill_technique="GPU Warp (Fast)",
                depth_blur_edge_threshold=20.0, depth_blur_strength=20.0, depth_map_blur=True,
                depth_blur_falloff=2.0, depth_blur_vert_smooth=6, batch_size=12)

CPU_FILLS = ['No fill', 'No fill - Reverse projection', 'Imperfect fill - Hybrid Edge', 'Fill - Naive',
             'Fill - Naive interpolating', 'Fill - Polylines Soft', 'Fill - Polylines Sharp']


def node_cases():
    cases = []

    def add(name, n, h, w, kind, **kw):
        spec = dict(name=name, n=n, h=h, w=w, kind=kind, seed=len(cases), channels=3,
                    scale255=False, black_box=False)
        for k in ("channels", "scale255", "black_box"):
            if k in kw:
                spec[k] = kw.pop(k)
        p = dict(DEFAULTS)
        p.update(kw)
        spec["params"] = p
        cases.append(spec)

    # every dropdown technique on the benchmark-like scene, big shifts for a small frame
    for i, fill in enumerate(CPU_FILLS + ['GPU Warp (Fast)']):
        add(f"scene_{i}", 2, 40, 96, "scene", fill_technique=fill, divergence=12.0)
    # BASELINE.json configs at fixture size
    add("cfg1_naive", 1, 48, 48, "scene", fill_technique='Fill - Naive', divergence=3.5, depth_map_blur=False)
    add("cfg2_polysharp", 1, 36, 128, "scene", fill_technique='Fill - Polylines Sharp', divergence=3.5)
    add("cfg3_hybrid", 3, 36, 96, "scene", fill_technique='Imperfect fill - Hybrid Edge', divergence=3.5)
    add("cfg4_gpuwarp_anaglyph", 3, 32, 96, "scene", fill_technique='GPU Warp (Fast)',
        modes='red-cyan-anaglyph', divergence=10.0, batch_size=2)
    add("cfg5_poly_balance", 1, 32, 128, "scene", fill_technique='Fill - Polylines Sharp',
        stereo_balance=0.5, divergence=4.5)
    # modes
    for i, mode in enumerate(["right-left", "top-bottom", "bottom-top", "red-cyan-anaglyph"]):
        add(f"mode_{i}", 1, 24, 64, "scene", fill_technique='Fill - Polylines Soft', modes=mode, divergence=8.0)
        add(f"mode_gw_{i}", 2, 24, 64, "scene", fill_technique='GPU Warp (Fast)', modes=mode, divergence=8.0)
    # depth classes x techniques
    for kind in ("noise", "quant", "flat", "steps", "card"):
        for fill in ('Fill - Polylines Sharp', 'Fill - Naive', 'Imperfect fill - Hybrid Edge',
                     'No fill - Reverse projection', 'GPU Warp (Fast)', 'Fill - Naive interpolating'):
            short = fill.split(' - ')[-1].replace(' ', '').replace('(Fast)', '')
            add(f"{kind}_{short}", 1, 32, 80, kind, fill_technique=fill, divergence=10.0)
    # parameter corners
    add("conv0", 1, 32, 80, "scene", fill_technique='Fill - Polylines Sharp', convergence_point=0.0, divergence=9.0)
    add("conv1", 1, 32, 80, "scene", fill_technique='Fill - Polylines Soft', convergence_point=1.0, divergence=9.0)
    add("sep_pos", 1, 32, 80, "scene", fill_technique='Fill - Naive', separation=2.5, divergence=9.0)
    add("sep_neg", 1, 32, 80, "scene", fill_technique='Fill - Polylines Sharp', separation=-3.0, divergence=9.0)
    add("bal_hi", 1, 32, 80, "scene", fill_technique='Fill - Polylines Sharp', stereo_balance=0.95, divergence=0.05)
    add("bal_lo_gw", 2, 32, 80, "scene", fill_technique='GPU Warp (Fast)', stereo_balance=-0.95, divergence=0.05)
    add("exp1", 1, 32, 80, "scene", fill_technique='Fill - Polylines Sharp', stereo_offset_exponent=1.0, divergence=6.0)
    add("exp07", 1, 32, 80, "scene", fill_technique='No fill', stereo_offset_exponent=0.7, divergence=6.0)
    add("exp13_gw", 1, 32, 80, "scene", fill_technique='GPU Warp (Fast)', stereo_offset_exponent=1.3, divergence=6.0)
    add("depth255", 1, 32, 80, "scene", fill_technique='Fill - Naive', scale255=True, divergence=9.0)
    add("depth255_gw", 2, 32, 80, "scene", fill_technique='GPU Warp (Fast)', scale255=True, divergence=9.0)
    add("depth1ch", 1, 32, 80, "scene", fill_technique='Fill - Polylines Soft', channels=1, divergence=9.0)
    add("blackbox", 1, 32, 80, "scene", fill_technique='Fill - Naive', black_box=True, divergence=9.0)
    add("blur_odd", 1, 32, 80, "steps", fill_technique='Fill - Polylines Sharp', depth_blur_strength=7.3,
        depth_blur_edge_threshold=4.0, depth_blur_falloff=1.0, depth_blur_vert_smooth=0, divergence=9.0)
    add("blur_wide", 1, 40, 96, "steps", fill_technique='No fill', depth_blur_strength=33.5,
        depth_blur_edge_threshold=2.0, depth_blur_falloff=0.5, depth_blur_vert_smooth=15, divergence=9.0)
    add("blur_off_gw", 2, 32, 80, "scene", fill_technique='GPU Warp (Fast)', depth_map_blur=False, divergence=9.0)
    add("batch_tail", 5, 24, 64, "scene", fill_technique='Fill - Naive', batch_size=2, divergence=9.0)
    add("batch_tail_gw", 5, 24, 64, "scene", fill_technique='GPU Warp (Fast)', batch_size=2, divergence=9.0)
    return cases


def crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def case_inputs(spec):
    img = syn.make_image(spec["n"], spec["h"], spec["w"], seed=spec["seed"], black_box=spec["black_box"])
    dep = syn.make_depth(spec["n"], spec.get("dh", spec["h"]), spec.get("dw", spec["w"]), spec["kind"],
                         seed=spec["seed"], channels=spec["channels"], scale255=spec["scale255"])
    return img, dep


def run_node(spec):
    """Runs the real node.  The reference's blur function is wrapped (not altered) so that the
    float32 blurred depth it produced is captured alongside the outputs: torch's conv2d
    summation order cannot be restated, so downstream integer parity is checked stage-wise by
    feeding THIS blurred depth to the implementation under test (parity protocol)."""
    Node = ref_loader.load_node_class()
    sig = sys.modules[Node.__module__].sig
    img, dep = case_inputs(spec)
    captured = []
    real_blur = sig.directional_motion_blur_gpu

    def spy(depth_tensor, *a, **k):
        L, R = real_blur(depth_tensor, *a, **k)
        captured.append((L.detach().cpu().numpy().copy(), R.detach().cpu().numpy().copy()))
        return L, R

    sig.directional_motion_blur_gpu = spy
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = Node().generate(torch.from_numpy(img), torch.from_numpy(dep), **spec["params"])
    finally:
        sig.directional_motion_blur_gpu = real_blur
    blur = None
    if captured:
        L = np.concatenate([c[0].reshape(-1, spec["h"], spec["w"]) for c in captured])
        R = np.concatenate([c[1].reshape(-1, spec["h"], spec["w"]) for c in captured])
        blur = (L.astype(np.float32), R.astype(np.float32))
    return img, dep, [o.numpy() for o in out], blur


def stage_cases():
    cases = []
    for kind in ("scene", "noise", "quant", "steps", "card"):
        for (s, thr, fo, v) in ((20, 20, 2.0, 6), (5.5, 3, 1.0, 0), (33.3, 1.0, 1.5, 15), (2.5, 20, 3.0, 1)):
            cases.append(dict(stage="blur", kind=kind, h=40, w=112, seed=len(cases),
                              strength=s, thr=thr, falloff=fo, vert=v))
    fills = ["none", "naive", "naive_interpolating", "polylines_soft", "polylines_sharp", "inverse", "hybrid_edge"]
    for kind in ("scene", "noise", "quant", "steps", "flat"):
        for (div, sep, expo, conv) in ((12.0, 0.0, 2.0, 0.5), (-9.0, 1.5, 1.0, 0.0), (7.0, -2.0, 0.7, 1.0)):
            for fill in fills:
                cases.append(dict(stage="warp", kind=kind, h=24, w=300, seed=len(cases), fill=fill,
                                  div=div, sep=sep, expo=expo, conv=conv))
    for kind in ("scene", "noise", "quant", "steps"):
        for (div, sep, expo, conv) in ((25.0, 0.0, 2.0, 0.5), (-18.0, 2.0, 1.0, 0.2), (11.0, -1.0, 0.5, 0.8)):
            cases.append(dict(stage="gpuwarp", kind=kind, h=24, w=160, seed=len(cases),
                              div_px=div, sep_px=sep, expo=expo, conv=conv))
    return cases


def stage_cases_post():
    """Second generation of stage fixtures (appended, existing files untouched): the post-fill variants."""
    cases = []
    for kind in ("scene", "noise", "quant", "steps", "flat"):
        for (div, sep, expo, conv) in ((12.0, 0.0, 2.0, 0.5), (-9.0, 1.5, 1.0, 0.0), (7.0, -2.0, 0.7, 1.0)):
            for fill in ("none_post", "inverse_post", "hybrid_edge_plus"):
                cases.append(dict(stage="warp", kind=kind, h=24, w=300, seed=500 + len(cases), fill=fill,
                                  div=div, sep=sep, expo=expo, conv=conv))
    return cases


def run_stage(spec):
    sig = ref_loader.load_sig()
    h, w = spec["h"], spec["w"]
    d = syn.make_depth(1, h, w, spec["kind"], seed=spec["seed"])[0, ..., 0]
    if spec["stage"] == "blur":
        d255 = d * np.float32(255)
        L, R = sig.directional_motion_blur_gpu(torch.from_numpy(d255), spec["strength"], spec["thr"],
                                               spec["strength"], falloff_exponent=spec["falloff"],
                                               vert_smooth_px=spec["vert"])
        return dict(crc=crc(d255), L=L.numpy(), R=R.numpy())
    if spec["stage"] == "warp":
        probe = syn.index_probe_image(h, w)
        d255 = d * np.float32(255)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = sig.apply_stereo_divergence(probe, d255, spec["div"], spec["sep"], spec["expo"],
                                              spec["fill"], spec["conv"])
        return dict(crc=crc(probe, d255), out=np.asarray(out))
    if spec["stage"] == "gpuwarp":
        img = syn.make_image(1, h, w, seed=spec["seed"])
        t = torch.from_numpy(img).permute(0, 3, 1, 2).contiguous()
        warped, mask = sig.forward_warp_gpu(t, torch.from_numpy(d[None]), spec["div_px"], spec["sep_px"],
                                            spec["expo"], spec["conv"])
        return dict(crc=crc(img, d), warped=warped[0].numpy(), mask=mask[0].numpy().astype(np.uint8))
    raise ValueError(spec["stage"])


def main_post():
    """python oracle/make_golden.py post  -- appends the post-fill stage fixtures to the manifest."""
    assert ref_loader.reference_available(), "needs /root/reference"
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        manifest = json.load(f)
    manifest["stage"] = [s for s in manifest["stage"] if not s["name"].startswith("post_")]
    for i, spec in enumerate(stage_cases_post()):
        rec = run_stage(spec)
        spec["name"] = f"post_{i:03d}"
        np.savez_compressed(os.path.join(GOLDEN, f"stage_{spec['name']}.npz"), **rec)
        manifest["stage"].append(spec)
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("post-fill stage cases:", len(stage_cases_post()))


# ---- N1 depth resize (GS:141-148 / GS:214-220) ---------------------------------------------------
# torch's CPU bilinear kernel contracts multiply-adds into FMAs in its AVX2/AVX512 builds (which ones
# depends on the tensor shape), so the reference's resized depth differs between machines by ~1 ulp of
# the source coordinate.  The fixtures are generated with ATEN_CPU_CAPABILITY=default (strict float32,
# what the oracle restates bit-exactly); `out_native` keeps this machine's default-dispatch result to
# bound the difference.
RESIZE_SHAPES = [(37, 53, 74, 106), (37, 53, 100, 91), (64, 64, 48, 40), (11, 7, 5, 3), (20, 30, 20, 45),
                 (30, 20, 45, 20), (96, 128, 48, 64), (1, 9, 4, 31), (9, 1, 31, 4), (135, 240, 270, 480)]


def resize_specs():
    kinds = ["scene", "noise", "card", "steps"]
    return [dict(name=f"{i:03d}", dh=dh, dw=dw, h=h, w=w, kind=kinds[i % len(kinds)], seed=300 + i)
            for i, (dh, dw, h, w) in enumerate(RESIZE_SHAPES)]


def node_resize_cases():
    def case(name, n, dh, dw, h, w, kind, channels, **params):
        p = dict(DEFAULTS)
        p.update(params)
        return dict(name=name, n=n, h=h, w=w, dh=dh, dw=dw, kind=kind, seed=400 + len(name), channels=channels,
                    scale255=False, black_box=False, params=p)
    return [
        case("rs_gw_up", 3, 20, 30, 48, 64, "scene", 3, fill_technique='GPU Warp (Fast)', depth_map_blur=False, batch_size=2),
        case("rs_gw_blur", 2, 27, 35, 48, 64, "scene", 1, fill_technique='GPU Warp (Fast)', depth_blur_strength=7.0),
        case("rs_naive_down", 2, 96, 128, 48, 64, "noise", 1, fill_technique='Fill - Naive', depth_map_blur=False),
        case("rs_sharp_up", 2, 37, 53, 64, 96, "scene", 3, fill_technique='Fill - Polylines Sharp', depth_map_blur=False),
        case("rs_soft_blur", 1, 37, 53, 64, 96, "scene", 3, fill_technique='Fill - Polylines Soft', depth_blur_strength=7.0,
             modes="top-bottom"),
        case("rs_hybrid_wide", 1, 48, 40, 48, 64, "card", 3, fill_technique='Imperfect fill - Hybrid Edge',
             depth_map_blur=False, modes="red-cyan-anaglyph"),
    ]


def _interp(gray, size):
    return torch.nn.functional.interpolate(torch.from_numpy(gray)[None, None], size=size, mode='bilinear',
                                           align_corners=False)[0, 0].numpy()


def main_resize():
    """python oracle/make_golden.py resize -- appends the depth-resize fixtures (stage + node) to the manifest."""
    assert ref_loader.reference_available(), "needs /root/reference"
    import subprocess
    import tempfile
    if os.environ.get("ATEN_CPU_CAPABILITY") != "default":
        # pass 1 (this machine's dispatch): the native results, then re-run strictly
        tmp = tempfile.mkdtemp()
        native = {}
        for spec in resize_specs():
            d = syn.make_depth(1, spec["dh"], spec["dw"], spec["kind"], seed=spec["seed"], channels=1)[0, ..., 0]
            native[spec["name"]] = _interp(d, (spec["h"], spec["w"]))
        np.savez(os.path.join(tmp, "native.npz"), **native)
        env = dict(os.environ, ATEN_CPU_CAPABILITY="default", CS_NATIVE_NPZ=os.path.join(tmp, "native.npz"))
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "resize"], env=env)
        return
    assert torch.backends.cpu.get_cpu_capability() == "DEFAULT"
    native = np.load(os.environ["CS_NATIVE_NPZ"])
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        manifest = json.load(f)
    manifest["resize"] = []
    for spec in resize_specs():
        d = syn.make_depth(1, spec["dh"], spec["dw"], spec["kind"], seed=spec["seed"], channels=1)[0, ..., 0]
        np.savez_compressed(os.path.join(GOLDEN, f"resize_{spec['name']}.npz"), crc=crc(d),
                            out_strict=_interp(d, (spec["h"], spec["w"])), out_native=native[spec["name"]])
        manifest["resize"].append(spec)
    manifest["node"] = [s for s in manifest["node"] if not s["name"].startswith("rs_")]
    for spec in node_resize_cases():
        img, dep, outs, blur = run_node(spec)
        np.savez_compressed(os.path.join(GOLDEN, f"node_{spec['name']}.npz"), **node_record(spec, img, dep, outs, blur))
        manifest["node"].append(spec)
        print("node", spec["name"], [o.shape for o in outs])
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("resize cases:", len(manifest["resize"]), "+", len(node_resize_cases()), "node cases")


def node_record(spec, img, dep, outs, blur):
    is_gw = spec["params"]["fill_technique"] == 'GPU Warp (Fast)'
    stereo, dl, dr, mask = outs
    rec = dict(crc=crc(img, dep))
    if is_gw:
        rec.update(stereo=stereo.astype(np.float32), depth_l=dl[..., 0].astype(np.float32),
                   depth_r=dr[..., 0].astype(np.float32), mask=(mask > 0).astype(np.uint8))
    else:
        q = lambda a: np.rint(a * 255.0).astype(np.uint8)
        assert np.array_equal(q(stereo).astype(np.float32) / np.float32(255), stereo)
        rec.update(stereo=q(stereo), depth_l=q(dl[..., 0]), depth_r=q(dr[..., 0]), mask=q(mask))
        assert np.array_equal(dl[..., 0], dl[..., 1]) and np.array_equal(dl[..., 0], dl[..., 2])
    if blur is not None:
        rec.update(blur_l=blur[0], blur_r=blur[1])
    return rec


DARK_CASES = [dict(name=f"{i:03d}", seed=sd, div=dv, sep=sp, fill=fl)
              for i, (sd, dv, sp, fl) in enumerate(
                  [(s_, d_, p_, 'naive_interpolating') for s_ in (0, 1, 2, 3, 4) for d_, p_ in ((6.0, 0.0), (-6.0, 1.0), (15.0, -2.0))]
                  + [(1, 6.0, 0.0, f_) for f_ in ('naive', 'polylines_soft', 'polylines_sharp', 'none_post', 'hybrid_edge_plus')])]


def main_dark():
    """python oracle/make_golden.py dark -- apply_stereo_divergence on dark images (found by oracle/fuzz_vs_reference.py:
    the interpolating fill's ramp arithmetic is float32, which only shows on a fraction of a percent of ramp values)."""
    assert ref_loader.reference_available(), "needs /root/reference"
    Node = ref_loader.load_node_class()
    sig = sys.modules[Node.__module__].sig
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        manifest = json.load(f)
    manifest["dark"] = []
    for spec in DARK_CASES:
        img, d = syn.dark_case(spec["seed"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = np.asarray(sig.apply_stereo_divergence(img.copy(), d.copy(), spec["div"], spec["sep"], 1.0, spec["fill"], 0.5))
        np.savez_compressed(os.path.join(GOLDEN, f"dark_{spec['name']}.npz"), crc=crc(img, d), out=out.astype(np.uint8))
        manifest["dark"].append(spec)
        print("dark", spec["name"], spec["fill"], flush=True)
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("dark cases:", len(DARK_CASES))


# ---- non-tensor inputs of create_stereoimages (numpy arrays / PIL images): SIG:1486-1496, scipy blur SIG:1346-1419 -------------
ARRAY_CASES = [
    dict(name="000", h=40, w=97, kind="scene", seed=1, fill="polylines_sharp", blur=True, s=20.0, thr=20.0, fo=2.0, v=6, div=8.0, sep=0.0, bal=0.0, expo=2.0, conv=0.5, modes=["left-right", "red-cyan-anaglyph"]),
    dict(name="001", h=33, w=64, kind="steps", seed=2, fill="naive", blur=True, s=7.3, thr=6.0, fo=1.0, v=0, div=6.0, sep=1.0, bal=0.3, expo=1.0, conv=0.0, modes=["top-bottom"]),
    dict(name="002", h=25, w=130, kind="noise", seed=3, fill="polylines_soft", blur=True, s=4.0, thr=2.0, fo=0.5, v=3, div=5.0, sep=-1.0, bal=-0.4, expo=2.0, conv=1.0, modes=["right-left"]),
    dict(name="003", h=60, w=80, kind="card", seed=4, fill="inverse", blur=True, s=21.0, thr=30.0, fo=3.0, v=15, div=10.0, sep=0.0, bal=0.0, expo=2.0, conv=0.5, modes=["left-right"]),
    dict(name="004", h=30, w=90, kind="quant", seed=5, fill="naive_interpolating", blur=False, s=0.0, thr=6.0, fo=1.0, v=0, div=7.0, sep=0.0, bal=0.0, expo=2.0, conv=0.5, modes=["bottom-top"]),
    dict(name="005", h=24, w=70, kind="scene", seed=6, fill="hybrid_edge", blur=True, s=9.0, thr=10.0, fo=2.0, v=2, div=9.0, sep=0.5, bal=0.0, expo=2.0, conv=0.3, modes=["left-right"]),
    dict(name="006", h=20, w=50, kind="scene", seed=7, fill="none", blur=True, s=33.0, thr=5.0, fo=2.0, v=2, div=4.0, sep=0.0, bal=0.95, expo=2.0, conv=0.5, modes=["left-right"]),
]


def array_inputs(spec):
    """uint8 RGB image and a float32 depth map on the 0..255 scale (what a PIL 'L' depth image gives), from seeds."""
    img = (syn.make_image(1, spec["h"], spec["w"], seed=spec["seed"])[0] * 255).astype(np.uint8)
    d = (syn.make_depth(1, spec["h"], spec["w"], spec["kind"], seed=spec["seed"], channels=1)[0, ..., 0] * np.float32(255))
    return img, d.astype(np.float32)


def main_arrays():
    """python oracle/make_golden.py arrays -- create_stereoimages with numpy inputs, and the scipy blur on its own."""
    assert ref_loader.reference_available(), "needs /root/reference"
    sig = ref_loader.load_sig()
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        manifest = json.load(f)
    manifest["arrays"] = []
    for spec in ARRAY_CASES:
        img, d = array_inputs(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = sig.create_stereoimages(img.copy(), d.copy(), spec["div"], spec["sep"], list(spec["modes"]), spec["bal"], spec["expo"],
                                          spec["fill"], spec["s"], spec["thr"], spec["blur"], True, spec["conv"], spec["fo"], spec["v"])
            rec = dict(crc=crc(img, d))
            for i, im in enumerate(out[0]):
                rec[f"stereo{i}"] = np.asarray(im)
            rec["depth_l"] = np.asarray(out[1])
            if spec["blur"]:
                rec["depth_r"] = np.asarray(out[2])
                bl, br = sig.directional_motion_blur(d.copy(), spec["s"], spec["thr"], spec["s"], falloff_exponent=spec["fo"],
                                                     vert_smooth_px=spec["v"])
                rec["blur_l"], rec["blur_r"] = bl.astype(np.float32), br.astype(np.float32)
        np.savez_compressed(os.path.join(GOLDEN, f"arrays_{spec['name']}.npz"), **rec)
        manifest["arrays"].append(spec)
        print("arrays", spec["name"], spec["fill"], flush=True)
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


def main():
    assert ref_loader.reference_available(), "needs /root/reference"
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {"node": [], "stage": []}
    for spec in node_cases():
        img, dep, outs, blur = run_node(spec)
        is_gw = spec["params"]["fill_technique"] == 'GPU Warp (Fast)'
        stereo, dl, dr, mask = outs
        rec = dict(crc=crc(img, dep))
        if is_gw:
            rec.update(stereo=stereo.astype(np.float32), depth_l=dl[..., 0].astype(np.float32),
                       depth_r=dr[..., 0].astype(np.float32), mask=(mask > 0).astype(np.uint8))
        else:
            q = lambda a: np.rint(a * 255.0).astype(np.uint8)
            assert np.array_equal(q(stereo).astype(np.float32) / np.float32(255), stereo)
            rec.update(stereo=q(stereo), depth_l=q(dl[..., 0]), depth_r=q(dr[..., 0]), mask=q(mask))
            assert np.array_equal(dl[..., 0], dl[..., 1]) and np.array_equal(dl[..., 0], dl[..., 2])
        if blur is not None:
            rec.update(blur_l=blur[0], blur_r=blur[1])
        np.savez_compressed(os.path.join(GOLDEN, f"node_{spec['name']}.npz"), **rec)
        manifest["node"].append(spec)
        print("node", spec["name"], [o.shape for o in outs])
    for i, spec in enumerate(stage_cases()):
        rec = run_stage(spec)
        spec["name"] = f"{spec['stage']}_{i:03d}"
        np.savez_compressed(os.path.join(GOLDEN, f"stage_{spec['name']}.npz"), **rec)
        manifest["stage"].append(spec)
    print("stage cases:", len(manifest["stage"]))
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "post":
        main_post()
    elif len(sys.argv) > 1 and sys.argv[1] == "resize":
        main_resize()
    elif len(sys.argv) > 1 and sys.argv[1] == "dark":
        main_dark()
    elif len(sys.argv) > 1 and sys.argv[1] == "arrays":
        main_arrays()
    else:
        main()
