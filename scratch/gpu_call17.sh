#!/bin/bash
# compute-sanitizer memcheck over the unit tests of the kernels added this session
set -u
OUT=gpurun_out; mkdir -p $OUT
SAN="compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9"
( timeout 280 $SAN python -m pytest tests/test_elem.py -m gpu -x -q -k "linear or gdfn or multi_m or dwconv_bwd_fused or conv_wgrad" 2>&1 | tail -25 ) > $OUT/c17_san_elem.log
tail -6 $OUT/c17_san_elem.log
( timeout 280 $SAN python -m pytest tests/test_gemm_pm.py -m gpu -x -q -k "tma or conv_fwd_dgrad" 2>&1 | tail -25 ) > $OUT/c17_san_pm.log
tail -6 $OUT/c17_san_pm.log
( timeout 200 $SAN python -m pytest tests/test_block.py -m gpu -x -q -k "test_block_fwd_bwd and 96-1" 2>&1 | tail -25 ) > $OUT/c17_san_block.log
tail -6 $OUT/c17_san_block.log
