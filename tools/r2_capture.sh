#!/bin/bash
# Round-2 evidence, one B200 (outputs under gpurun_out/, copied to profiles/ by hand):  bash tools/r2_capture.sh
set -u
B="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-images 0 --no-extra"
$B > gpurun_out/r2_cap_bench.json 2> gpurun_out/r2_cap_bench.err; echo "bench (no ncu) rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:haar_icon_tma2 -s 3 -c 2 -o gpurun_out/r2_icon $B > /dev/null 2>&1
ncu -i gpurun_out/r2_icon.ncu-rep --page raw --csv > gpurun_out/r2_icon_tma2_full_raw.csv
python tools/profile_epilogue.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:resize_area_rows -s 1 -c 4 -o gpurun_out/r2_epi python tools/profile_epilogue.py > /dev/null 2>&1
ncu -i gpurun_out/r2_epi.ncu-rep --page raw --csv > gpurun_out/r2_epilogue_rows_full_raw.csv
python tools/bench_rows.py deep > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tail -c 4 -o gpurun_out/r2_tailk python tools/bench_rows.py deep > /dev/null 2>&1
ncu -i gpurun_out/r2_tailk.ncu-rep --page raw --csv > gpurun_out/r2_rows_tail_full_raw.csv
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:"rows_kernel|tail|tma2" -c 400 --csv --log-file gpurun_out/r2_launches_deep.csv python tools/bench_rows.py deep > /dev/null 2>&1
python tools/profile_subband.py 6 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:patch -s 2 -c 2 -o gpurun_out/r2_sub6 python tools/profile_subband.py 6 > /dev/null 2>&1
ncu -i gpurun_out/r2_sub6.ncu-rep --page raw --csv > gpurun_out/r2_subband_patch_d6_full_raw.csv
WICCA_ROWS_SUBBAND_DEPTHS=1,3,6 python tools/bench_rows.py > gpurun_out/r2_rows.jsonl 2> gpurun_out/r2_rows.err; echo "rows rc=$?"; tail -n 3 gpurun_out/r2_rows.err
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | grep r2_ | tail -20
