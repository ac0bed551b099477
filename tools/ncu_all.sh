#!/bin/bash
# One `ncu --set full` capture of every kernel of the library (run under gpurun, one GPU): reports land in gpurun_out/r02_ncu_*.ncu-rep
# kbench.py issues 6 calls of the kernel sequence; the first 4 are skipped, the 5th is captured.
cap() {  # name, kernels per call, kbench args...
  name=$1; per=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_ -s $((4 * per)) -c $per \
      -o gpurun_out/r02_ncu_$name python tools/kbench.py --frames 4 --steps 1 "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
cap polylines_sharp 7 --fill polylines_sharp
cap polylines_soft 7 --fill polylines_soft
cap naive 6 --fill naive
cap naive_interp_noblur 5 --fill naive_interpolating --no-blur
cap inverse 6 --fill inverse
cap hybrid 7 --fill hybrid_edge
cap gpuwarp 6 --fill gpu_warp
cap gpuwarp4k 6 --fill gpu_warp --width 3840 --height 2160 --mode red-cyan-anaglyph --divergence 10
cap poly8k 7 --fill polylines_sharp --width 7680 --height 3840 --frames 2 --balance 0.5
