#!/bin/bash
# Round 2, seventh GPU call (1 GPU): multigrid fast paths (timing + launch list), then the whole suite and the bench
set -u
O=gpurun_out/r2s7
mkdir -p "$O"
for n in 512 128; do timeout 600 python tools/profile_mg.py $n solve >> "$O/mg_solve.log" 2>&1; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file "$O/launches_mg512.csv" python tools/profile_mg.py 512 cycle > "$O/ncu_mg512.log" 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$O/pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$O/summary.txt"
timeout 900 python bench.py > "$O/bench_n1.json" 2> "$O/bench_n1.err"
echo "bench rc=$?" >> "$O/summary.txt"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > "$O/bench_reference_n1.json" 2> "$O/bench_reference_n1.err"
python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1
echo "smoke rc=$?" >> "$O/summary.txt"
ls -la "$O" > "$O/listing.txt"
