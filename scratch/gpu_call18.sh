#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the CUDA-core kernels added this session
set -u
OUT=gpurun_out; mkdir -p $OUT
SAN="compute-sanitizer --tool racecheck --print-limit 10 --error-exitcode 9"
( timeout 280 $SAN python -m pytest tests/test_elem.py -m gpu -x -q -k "linear_kernels and 3-50 or gdfn_mid_bwd_fused and 1-4 or dwconv_bwd_fused and 3-37" 2>&1 | tail -25 ) > $OUT/c18_race_elem.log
tail -8 $OUT/c18_race_elem.log
( timeout 280 $SAN python -m pytest tests/test_block.py -m gpu -x -q -k "test_block_fwd_bwd and True-96-2" 2>&1 | tail -30 ) > $OUT/c18_race_block.log
tail -12 $OUT/c18_race_block.log
