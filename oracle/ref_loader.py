"""Load the UNMODIFIED reference (ComfyStereo) from /root/reference for oracle pinning.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (to generate the committed
fixtures under tests/golden/) and by tests that cross-check the C oracle against the
live reference when /root/reference exists (it does not exist on the GPU box).
Nothing in the product package imports this file.

The reference package needs `comfy.utils.ProgressBar` (GenerateStereo.py:27); ComfyUI is
not installed, so a 6-line stub is registered before import.  The tree is read-only, so
bytecode writing is disabled.
"""
import importlib
import importlib.util
import os
import sys
import types

REFERENCE_DIR = os.environ.get("COMFYSTEREO_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "stereoimage_generation.py"))


def _install_comfy_stub():
    if "comfy.utils" in sys.modules:
        return
    comfy = types.ModuleType("comfy")
    utils = types.ModuleType("comfy.utils")

    class ProgressBar:  # same surface the node uses: ProgressBar(total).update(k)
        def __init__(self, total):
            self.total, self.done = total, 0

        def update(self, k):
            self.done += k

    utils.ProgressBar = ProgressBar
    comfy.utils = utils
    sys.modules["comfy"] = comfy
    sys.modules["comfy.utils"] = utils


_cache = {}


def load_sig():
    """The hot-path module alone (stereoimage_generation.py)."""
    if "sig" not in _cache:
        sys.dont_write_bytecode = True
        spec = importlib.util.spec_from_file_location(
            "_ref_sig", os.path.join(REFERENCE_DIR, "stereoimage_generation.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache["sig"] = mod
    return _cache["sig"]


def load_node_class():
    """The real StereoImageNode class, imported as a package with the comfy stub."""
    if "node" not in _cache:
        sys.dont_write_bytecode = True
        _install_comfy_stub()
        parent = os.path.dirname(REFERENCE_DIR.rstrip("/"))
        name = os.path.basename(REFERENCE_DIR.rstrip("/"))
        if parent not in sys.path:
            sys.path.insert(0, parent)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            pkg = importlib.import_module(name)
        _cache["node"] = pkg.NODE_CLASS_MAPPINGS["StereoImageNode"]
        _cache["pkg"] = pkg
    return _cache["node"]
