#!/bin/bash
# Device-resident throughput of the five BASELINE.json configs (+ the other techniques at 1080p) -> gpurun_out/configs.jsonl
out=gpurun_out/configs.jsonl; : > $out
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --strong-frames 0 "$@" 2>/dev/null | tail -1 >> $out; }
run --height 512 --width 512 --frames 64 --fill "Fill - Naive"
run --fill "Fill - Polylines Sharp"
run --fill "Imperfect fill - Hybrid Edge"
run --height 2160 --width 3840 --frames 8 --fill "GPU Warp (Fast)" --mode red-cyan-anaglyph --divergence 10
run --height 3840 --width 7680 --frames 2 --fill "Fill - Polylines Sharp" --balance 0.5
for f in "GPU Warp (Fast)" "No fill" "No fill - Reverse projection" "Fill - Naive" "Fill - Naive interpolating" "Fill - Polylines Soft" "Fill - Post-fill" "Fill - Reverse projection with Post-fill" "Fill - Hybrid Edge with fill"; do run --fill "$f"; done
COMFYSTEREO_GPU_WARP=mesh run --fill "GPU Warp (Fast)"                                             # forward_warp_mesh
COMFYSTEREO_GPU_WARP=mesh run --height 2160 --width 3840 --frames 8 --fill "GPU Warp (Fast)" --mode red-cyan-anaglyph --divergence 10
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    d=json.loads(l); c=d['config']['workload'][:72]
    print(f"{c:74s} frames/step {d['config']['frames_per_gpu_per_step']:3d}  {d['value']:9.1f} fps  {d['mpix_per_s']:9.0f} Mpix/s  path-roofline {100*d['roofline']['path']['frac']:5.1f}%  top {d['roofline']['kernel']} {100*d['roofline']['share_of_step']:.0f}%")
PY
