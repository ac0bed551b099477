"""Dev telemetry: how many rows of a benchmark frame fall back to the sequential replay."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comfystereo_b200 import engine, _lib, synthetic as syn
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
div = float(sys.argv[3]) if len(sys.argv) > 3 else 3.5
bal = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
img = torch.from_numpy(syn.make_image(1, h, w, seed=0)).cuda()
dep = torch.from_numpy(syn.make_depth(1, h, w, "scene", seed=0)).cuda()
p = engine.make_params("polylines_sharp", "left-right", div, 0, bal, 0.5, 2.0, True, 20, 20, 2.0, 6)
for _ in range(2):
    engine.stereo_batch_device(img, dep, p, chunk=1)
torch.cuda.synchronize()
ws = engine._workspaces[0]
st, fl = ctypes.c_int(), ctypes.c_int()
_lib.check(_lib.lib().cs_polylines_status(ctypes.byref(p), 1, h, w, ws.data_ptr(), ctypes.byref(st), ctypes.byref(fl)))
print("status", st.value, "flagged rows", fl.value, "of", 2 * h)
