#!/bin/bash
# quick GPU iteration loop: stage times (prep / ecc / warp per frame) and frames/s on 16 frames of the 4K stack
# usage: scripts/quick_bench.sh [extra bench.py args, e.g. --motion 2]
python bench.py --frames 16 --steps 3 --warmup 2 --skip-cpu --skip-e2e "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('frames/s %.0f  K2 us/launch %.1f  prep %.1f us  warp %.1f us  ecc loop %.1f us/frame (%.2f it)  err %.3f px' % (
  d['value'], d['roofline']['us_per_launch'], 1e3*d['stages']['prep']['ms_per_frame'], 1e3*d['stages']['warp_accumulate']['ms_per_frame'],
  1e3*d['stages']['ecc_loop']['ms_per_frame'], d['stages']['ecc_loop']['iterations_per_frame'], d['check']['max_corner_error_vs_ground_truth_px']))
"
