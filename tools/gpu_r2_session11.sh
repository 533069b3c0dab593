#!/bin/bash
set -u
O=gpurun_out/r2s11
mkdir -p "$O"
timeout 600 python tools/profile_forms.py 192 256 320 384 448 512 > "$O/forms_by_size.json" 2> "$O/forms.err"
timeout 300 python -m pytest tests/test_zz_b_cg_variant2.py tests/test_zz_c_single_reduction.py tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider > "$O/pytest_forms.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
