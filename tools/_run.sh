timeout 600 python -m pytest tests -m gpu -x -q -k "warp_stage or polylines or random_stage" 2>&1 | tail -3
for cfg in "4 0" "16 0"; do set -- $cfg; COMFYSTEREO_POLY_NW=$1 COMFYSTEREO_POLY_OCC=$2 timeout 120 python tools/kbench.py --tag nw$1/occ$2 --steps 10; done 2>&1 | grep -v Warn
