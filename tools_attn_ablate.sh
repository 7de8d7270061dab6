#!/bin/bash
# Timing-only ablations of the attention kernel (GPU box). Results are numerically wrong for ABL != 0.
for cfg in "0 3" "0 2" "1 3" "2 3" "4 3" "8 3" "16 3" "6 3" "14 3" "31 3"; do
  set -- $cfg
  AIGV_NVCC_EXTRA="-DAIGV_ABL=$1 -DAIGV_ATTN_MIN_CTAS=$2" python -c "
from aigv_assessor_b200 import _lib
_lib.build(force=True)" && echo -n "ABL=$1 CTAS=$2: " && python tools_op_bench.py attn 10
done
python -c "
from aigv_assessor_b200 import _lib
_lib.build(force=True)"
