#!/bin/bash
# ncu evidence for profiles/r2_summary.md (run on the GPU box through gpurun; numbers printed under ncu are never bench values)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
# 1. launch list of one timed step of the default bench workload (K+O, 10,000 assemblies): kernel shares
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kb_|gotoh|type_translate|DeviceRadixSort|DeviceSelect|DeviceScan|DeviceRunLength" --csv \
    --log-file $O/launches_r2.csv python bench.py --steps 1 --warmup 1 --e2e-asm 512 --e2e-ascii-asm 0 --no-cpu-baseline > $O/launches_r2.log 2>&1
# 2. the roofline kernel: one scan launch over the 10,000-assembly batch, full set (dram bytes = roofline.traffic)
if [ "${SKIP_SCAN:-0}" != 1 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kb_scan_kernel -s 1 -c 1 -o $O/r2_scan \
    python bench.py --steps 1 --warmup 1 --e2e-asm 512 --e2e-ascii-asm 0 --no-cpu-baseline > $O/r2_scan.log 2>&1
ncu -i $O/r2_scan.ncu-rep --page raw --csv > $O/r2_scan_raw.csv
fi
# 3. the DP kernels (1000 assemblies)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kb_rows16_kernel|kb_band16_kernel" -s 4 -c 4 -o $O/r2_dp \
    python bench.py --n-asm 1000 --steps 1 --warmup 1 --e2e-asm 512 --e2e-ascii-asm 0 --no-cpu-baseline > $O/r2_dp.log 2>&1
ncu -i $O/r2_dp.ncu-rep --page raw --csv > $O/r2_dp_raw.csv
rm -f $O/r2_scan.ncu-rep $O/r2_dp.ncu-rep
