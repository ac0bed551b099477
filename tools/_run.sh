timeout 600 python -m pytest tests -m gpu -x -q -k "warp_stage or polylines or random_stage" 2>&1 | tail -3
for nw in 4 8 16; do for occ in 0 1; do COMFYSTEREO_POLY_NW=$nw COMFYSTEREO_POLY_OCC=$occ timeout 120 python tools/kbench.py --tag nw$nw/occ$occ --steps 10; done; done 2>&1 | grep -v Warn
COMFYSTEREO_POLY_NW=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polylines -s 4 -c 1 -o gpurun_out/r02c_poly_nw16 python tools/kbench.py --frames 4 --steps 2 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
