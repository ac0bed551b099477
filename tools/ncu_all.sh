#!/bin/bash
# One `ncu --set full` capture of every kernel of the library (run under gpurun, one GPU).  The reports are summarised on the
# box (gpurun_out/ is limited to 64 MiB): gpurun_out/r02_ncu_kernels.md (one row per kernel) and one summary text per capture;
# only the Polylines report itself is kept.
# kbench.py issues 6 calls of the kernel sequence; the first 4 are skipped, the 5th is captured.
: > gpurun_out/r02_ncu_kernels_rows.md
cap() {  # name, kernels per call, kbench args...
  name=$1; per=$2; shift 2
  src=""; [ "$name" = polylines_sharp ] && src="--import-source on"
  timeout 600 ncu --set full --clock-control none $src -k regex:^k_ -s $((4 * per)) -c $per \
      -o gpurun_out/r02_ncu_$name python tools/kbench.py --frames 4 --steps 1 "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  python tools/ncu_table.py gpurun_out/r02_ncu_$name.ncu-rep | tail -n +3 >> gpurun_out/r02_ncu_kernels_rows.md
  [ "$name" = polylines_sharp ] || rm -f gpurun_out/r02_ncu_$name.ncu-rep
  rm -f gpurun_out/ncu_$name.log
}
cap polylines_sharp 6 --fill polylines_sharp
cap polylines_soft 6 --fill polylines_soft
cap naive 5 --fill naive
cap naive_interp_noblur 4 --fill naive_interpolating --no-blur
cap inverse_anaglyph 6 --fill inverse --mode red-cyan-anaglyph
cap hybrid 6 --fill hybrid_edge
cap gpuwarp 6 --fill gpu_warp
cap meshwarp 7 --fill gpu_warp_mesh
cap gpuwarp4k 6 --fill gpu_warp --width 3840 --height 2160 --mode red-cyan-anaglyph --divergence 10
cap poly8k 6 --fill polylines_sharp --width 7680 --height 3840 --frames 2 --balance 0.5
python tools/ncu_table.py /dev/null 2>/dev/null | head -2 > gpurun_out/r02_ncu_kernels.md
cat gpurun_out/r02_ncu_kernels_rows.md >> gpurun_out/r02_ncu_kernels.md; rm -f gpurun_out/r02_ncu_kernels_rows.md
du -sh gpurun_out
