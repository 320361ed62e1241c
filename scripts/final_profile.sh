#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): tests, bench (both arms), ncu launch list and full captures.
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/final_gputests.log
(timeout 900 python bench.py 2>&1 | tail -1) > gpurun_out/final_bench.json
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1) > gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kb_|DeviceRadixSort|DeviceSelect|DeviceScan" --csv \
    --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 1 --warmup 1 --e2e-asm 8 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kb_scan_kernel -s 1 -c 1 -o gpurun_out/prof_scan_final \
    python bench.py --steps 1 --warmup 1 --e2e-asm 8 --no-cpu-baseline > gpurun_out/ncu_scan_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"kb_rows_kernel|kb_band_kernel" -s 2 -c 2 -o gpurun_out/prof_dp_final \
    python bench.py --n-asm 250 --steps 1 --warmup 1 --e2e-asm 8 --no-cpu-baseline > gpurun_out/ncu_dp_final.log 2>&1
cat gpurun_out/final_gputests.log
cut -c1-400 gpurun_out/final_bench.json
cut -c1-300 gpurun_out/final_bench_reference.json
