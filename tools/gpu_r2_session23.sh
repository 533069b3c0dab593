#!/bin/bash
# Round 2, session 23 (the last GPU seconds): multigrid GPU tests after the attribute / scratch change, smoke
set -u
O=gpurun_out/r2s23
mkdir -p "$O"
timeout 45 python -m pytest tests/test_zz_multigrid.py tests/test_zz_b_cg_variant2.py -m gpu -q -p no:cacheprovider > "$O/pytest_mg.log" 2>&1
echo "pytest rc=$?" > "$O/summary.txt"
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1
echo "smoke rc=$?" >> "$O/summary.txt"
