#!/bin/bash
set -u
O=gpurun_out/r2s13
mkdir -p "$O"
timeout 400 python tools/sweep_variant2.py 512 > "$O/sweep_variant2_512.log" 2> "$O/sweep.err"
