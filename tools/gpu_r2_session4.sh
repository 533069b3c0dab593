#!/bin/bash
# Round 2, fourth GPU call (1 GPU): the whole GPU suite again, advection occupancy variants + ncu after the remap,
# (record of what ran: the advect_occ and mg_inorder keys measured here lost and were removed afterwards)
# multigrid launch order
set -u
O=gpurun_out/r2s4
mkdir -p "$O"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
for occ in 5 6 8; do
python - $occ >> "$O/advect_occ.json" 2>> "$O/advect_occ.err" <<'PY'
import json, sys
sys.path.insert(0, ".")
from cajitafluids_b200 import Solver, default_config
occ = int(sys.argv[1])
for n, tile in ((512, 0), (512, 1), (256, 0), (128, 0)):
    cfg = default_config(3, n, box=n / 512.0)
    cfg.cg_fixed_iters = 5
    s = Solver(cfg)
    s.set_tuning("advect_occ", occ)
    s.set_tuning("advect_tile", tile)
    s.setup(); s.step(); s.reset_stats()
    for _ in range(3):
        s.step()
    print(json.dumps({"cells": n, "advect_occ": occ, "advect_tile": tile, "ms_advect": s.stats()["ms_advect"] / 3}), flush=True)
    s.close()
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o "$O/advect512_remap_full" \
    python tools/profile_advect_target.py 512 0 > "$O/ncu_advect.log" 2>&1
# multigrid: launch order A/B (solve timings), launch list with the in-order form
python - > "$O/mg_inorder.log" 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
from cajitafluids_b200 import Solver, default_config
for n in (512, 256, 128):
    for inorder in (0, 1):
        cfg = default_config(3, n); cfg.cg_max_iter = 20000
        s = Solver(cfg); s.set_preconditioner("mg"); s.set_tuning("mg_inorder", inorder)
        s.add_inputs(); s.build_rhs(); s.pcg_solve(); s.reset_stats()
        t0 = time.perf_counter(); it, res = s.pcg_solve(); dt = time.perf_counter() - t0
        print(f"n={n} mg_inorder={inorder}: {it} iterations, |r|={res:.3e}, wall {dt*1e3:.2f} ms, {dt*1e3/it:.3f} ms/iteration", flush=True)
        s.close()
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file "$O/launches_mg512_inorder.csv" \
    python - > "$O/ncu_mg512.log" 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
from cajitafluids_b200 import Solver, default_config
cfg = default_config(3, 512); cfg.cg_fixed_iters = 3
s = Solver(cfg); s.set_preconditioner("mg"); s.set_tuning("mg_inorder", 1)
s.add_inputs(); s.build_rhs(); print(s.pcg_solve())
PY
# whole timestep numbers with the current advection (128^3 default problem, 512^3 with body force come with bench)
timeout 900 python bench.py --no-cpu-baseline --no-probe > "$O/bench_n1_quick.json" 2> "$O/bench_n1_quick.err"
ls -la "$O" > "$O/listing.txt"
