"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header declares,
parameter marshalling follows the reference's python arithmetic, the node keeps the reference's schema,
frame sharding covers the batch in order (also across 2 gloo ranks), and nothing in the product falls
back to a CPU implementation."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT
from comfystereo_b200 import _lib, engine
import comfystereo_b200 as pkg


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "comfystereo_b200.h")).read()
    return sorted(set(re.findall(r"CS_API [^;(]*?\b(cs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} is declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table binds exactly the header's surface
    assert _lib.lib().cs_abi_version() == 4


def test_graft_entry_build():
    """The driver's build check: compiles (no-op when up to date), loads the library, checks the ABI version."""
    import __graft_entry__ as g
    g.build()


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib().cs_device_check() == -3          # CS_ERR_DEVICE
    from comfystereo_b200 import stereoimage_generation as sig
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sig.create_stereoimages(torch.rand(3, 8, 16), torch.rand(8, 16), 3.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.StereoImageNode().generate(torch.rand(1, 8, 16, 3), torch.rand(1, 8, 16, 3), 3.5, 0.0, "left-right", 0.0,
                                       0.5, 2.0, "Fill - Naive", 20.0, 20.0, True)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under comfystereo_b200/ may import, link or execute it
    (comments may mention it)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "comfystereo_b200")):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            for line in open(os.path.join(dirpath, f)):
                code = line.split("//")[0].split("#")[0] if f.endswith((".cu", ".cuh", ".h")) else line.split("#")[0]
                assert "import oracle" not in code and "stereo_oracle" not in code and "libstereo_oracle" not in code, \
                    f"{f}: {line.strip()}"


def test_compact_transport_policy(monkeypatch):
    """The host path compacts the depth outputs / mask on the bus only when it may use >= 8 host threads per pipeline
    (hardware threads / LOCAL_WORLD_SIZE); COMFYSTEREO_COMPACT_D2H overrides.  Pure host logic: no GPU needed."""
    lib = _lib.lib()
    hw = len(os.sched_getaffinity(0))     # what std::thread::hardware_concurrency() reports
    monkeypatch.delenv("COMFYSTEREO_COMPACT_D2H", raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert lib.cs_host_compact_enabled() == (1 if hw >= 8 else 0)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", str(max(hw, 1)))      # one thread per rank: never
    assert lib.cs_host_compact_enabled() == 0
    monkeypatch.setenv("COMFYSTEREO_COMPACT_D2H", "1")
    assert lib.cs_host_compact_enabled() == 1
    monkeypatch.setenv("COMFYSTEREO_COMPACT_D2H", "0")
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert lib.cs_host_compact_enabled() == 0


def test_params_follow_python_rounding():
    # bs = int(round(s)) is banker's rounding, R = int(s) truncates (SIG:1208-1209, quirk Q11)
    for s, bs, r in ((2.5, 2, 2), (3.5, 4, 3), (20.5, 20, 20), (21.5, 22, 21), (20.0, 20, 20), (0.7, 1, 0)):
        p = engine.make_params('naive', 'left-right', 3.5, depth_blur=True, depth_blur_strength=s)
        assert (p.blur_enabled, p.blur_box, p.blur_radius) == (1, bs, r)
    with pytest.raises(RuntimeError, match="kernel size should be greater than zero"):
        engine.make_params('naive', 'left-right', 3.5, depth_blur=True, depth_blur_strength=0.4)
    assert engine.make_params('naive', 'left-right', 3.5, depth_blur=True, depth_blur_strength=0.0).blur_enabled == 0
    assert engine.make_params('naive', 'left-right', 3.5, depth_blur=False, depth_blur_strength=9.0).blur_enabled == 0
    with pytest.raises(ValueError, match="Unknown mode"):
        engine.make_params('naive', 'diagonal', 3.5)


def test_output_shapes_follow_the_reference():
    mk = lambda fill, mode: engine.make_params(fill, mode, 3.5)
    assert engine.output_shapes(mk('naive', 'left-right'), 2, 10, 16) == ((2, 10, 32, 3), (2, 10, 16, 3), (2, 10, 32))
    assert engine.output_shapes(mk('naive', 'top-bottom'), 2, 10, 16) == ((2, 20, 16, 3), (2, 10, 16, 3), (2, 20, 16))
    assert engine.output_shapes(mk('naive', 'red-cyan-anaglyph'), 1, 10, 16) == ((1, 10, 16, 3), (1, 10, 16, 3), (1, 10, 16))
    # GPU Warp: the mask keeps the single-eye shape even for SBS
    assert engine.output_shapes(mk('gpu_warp', 'left-right'), 2, 10, 16) == ((2, 10, 32, 3), (2, 10, 16, 3), (2, 10, 16))


def test_workspace_grows_with_chunk_and_technique():
    lib = _lib.lib()
    p0 = engine.make_params('naive', 'left-right', 3.5)
    p1 = engine.make_params('polylines_sharp', 'left-right', 3.5, depth_blur=True, depth_blur_strength=20)
    a = lib.cs_workspace_bytes(ctypes.byref(p0), 1, 1080, 1920)
    b = lib.cs_workspace_bytes(ctypes.byref(p1), 1, 1080, 1920)
    c = lib.cs_workspace_bytes(ctypes.byref(p1), 4, 1080, 1920)
    assert 0 < a < b < c and c <= 4 * b


def test_node_schema_matches_reference():
    node = pkg.NODE_CLASS_MAPPINGS["StereoImageNode"]
    assert pkg.NODE_DISPLAY_NAME_MAPPINGS == {"StereoImageNode": "Stereo Image Node"}
    assert node.RETURN_TYPES == ("IMAGE", "IMAGE", "IMAGE", "MASK")
    assert node.RETURN_NAMES == ("stereoscope", "blurred_depthmap_left", "blurred_depthmap_right", "no_fill_imperfect_mask")
    assert node.FUNCTION == "generate" and not hasattr(node, "CATEGORY")
    it = node.INPUT_TYPES()
    assert list(it["required"]) == ["image", "depth_map", "modes", "fill_technique"]
    assert it["required"]["modes"][0] == ["left-right", "right-left", "top-bottom", "bottom-top", "red-cyan-anaglyph"]
    assert it["required"]["fill_technique"][0][0] == 'GPU Warp (Fast)' and len(it["required"]["fill_technique"][0]) == 8
    assert it["required"]["fill_technique"][1]["default"] == "GPU Warp (Fast)"
    want = {"divergence": (4.5, 0.05, 15), "separation": (0, -5, 5), "stereo_balance": (0, -0.95, 0.95),
            "convergence_point": (0.5, 0.0, 1.0), "stereo_offset_exponent": (2, 0.1, 2),
            "depth_blur_edge_threshold": (20, 0.1, 60), "depth_blur_strength": (20, 0.1, 200),
            "depth_blur_falloff": (2.0, 0.1, 4.0), "depth_blur_vert_smooth": (6, 0, 15), "batch_size": (12, 1, 64)}
    assert list(it["optional"]) == ["divergence", "separation", "stereo_balance", "convergence_point",
                                    "stereo_offset_exponent", "depth_map_blur", "depth_blur_edge_threshold",
                                    "depth_blur_strength", "depth_blur_falloff", "depth_blur_vert_smooth", "batch_size"]
    for k, (d, lo, hi) in want.items():
        o = it["optional"][k][1]
        assert (o["default"], o["min"], o["max"]) == (d, lo, hi), k
    assert it["optional"]["depth_map_blur"] == ("BOOLEAN", it["optional"]["depth_map_blur"][1]) and \
        it["optional"]["depth_map_blur"][1]["default"] is True
    import inspect
    sig = inspect.signature(node.generate)
    assert list(sig.parameters)[1:] == ["image", "depth_map", "divergence", "separation", "modes", "stereo_balance",
                                        "convergence_point", "stereo_offset_exponent", "fill_technique",
                                        "depth_blur_edge_threshold", "depth_blur_strength", "depth_map_blur",
                                        "depth_blur_falloff", "depth_blur_vert_smooth", "batch_size"]
    assert (sig.parameters["depth_blur_falloff"].default, sig.parameters["depth_blur_vert_smooth"].default,
            sig.parameters["batch_size"].default) == (1.0, 0, 4)


def test_function_signatures_match_reference():
    import inspect
    from comfystereo_b200 import stereoimage_generation as sig
    a = inspect.signature(sig.create_stereoimages)
    assert list(a.parameters) == ["original_image", "depthmap", "divergence", "separation", "modes", "stereo_balance",
                                  "stereo_offset_exponent", "fill_technique", "depth_blur_strength",
                                  "depth_blur_edge_threshold", "direction_aware_depth_blur", "return_modified_depth",
                                  "convergence_point", "depth_blur_falloff", "depth_blur_vert_smooth"]
    assert a.parameters["fill_technique"].default == 'polylines_sharp' and a.parameters["stereo_offset_exponent"].default == 1.0
    b = inspect.signature(sig.create_stereoimages_gpu)
    assert list(b.parameters) == ["image_tensor", "depth_tensor", "divergence", "separation", "modes", "stereo_balance",
                                  "stereo_offset_exponent", "convergence_point", "depth_blur_strength",
                                  "depth_blur_edge_threshold", "direction_aware_depth_blur", "depth_blur_falloff",
                                  "depth_blur_vert_smooth"]


def test_shard_ranges_cover_in_order():
    for n in (1, 7, 60, 120, 300, 301):
        for world in (1, 2, 4, 8):
            spans = [engine.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            for group in (1, 4, 12):
                g = [engine.group_aligned_shard_range(n, r, world, group) for r in range(world)]
                assert g[0][0] == 0 and g[-1][1] == n and all(a[1] == b[0] for a, b in zip(g, g[1:]))
                assert all(lo % group == 0 for lo, hi in g if hi > lo)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from comfystereo_b200 import engine
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
lo, hi = engine.shard_range(n, rank, world)
# every rank "processes" its frames (tags them with the frame id) -- no collective on the data path;
# only the in-order output assembly gathers, as the north star allows
mine = torch.arange(lo, hi, dtype=torch.int64)
sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([hi - lo]))
parts = [torch.zeros(int(s), dtype=torch.int64) for s in sizes]
dist.all_gather(parts, mine) if len(set(int(s) for s in sizes)) == 1 else dist.all_gather_object(parts, mine)
out = torch.cat([torch.as_tensor(p) for p in parts])
assert torch.equal(out, torch.arange(n)), out
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)      # the bench's max-over-ranks timing reduction
assert t.item() == world
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), ROOT],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_k_over_255_table_is_the_ieee_quotient():
    """cs_internal.cuh's literal table kQ255 (what the kernels dequantise with) must be float32(k) / float32(255) bit for bit."""
    import re
    import numpy as np
    src = open(os.path.join(ROOT, "comfystereo_b200", "csrc", "cs_internal.cuh")).read()
    body = src[src.index("kQ255[256] = {") + len("kQ255[256] = {"):]
    body = body[:body.index("};")]
    vals = [float.fromhex(t.rstrip("f")) for t in re.findall(r"0x[0-9a-fA-F.]+p[-+]?\d+f", body)]
    assert len(vals) == 256
    want = np.arange(256, dtype=np.float32) / np.float32(255)
    assert np.array_equal(np.array(vals, np.float64).astype(np.float32), want)
    assert np.array_equal(np.array(vals, np.float64), want.astype(np.float64))      # exactly representable as float32
