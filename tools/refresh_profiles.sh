#!/bin/bash
# Recapture the ncu evidence of the secondary kernels under gpurun_out/ (copied to profiles/ by hand):
#   bash tools/refresh_profiles.sh
ncu --set full --clock-control none --import-source on -k regex:patch -s 2 -c 2 -o gpurun_out/sub6 python tools/profile_subband.py 6 > /dev/null 2>&1; ncu -i gpurun_out/sub6.ncu-rep --page raw --csv > gpurun_out/sub6_raw.csv
ncu --set full --clock-control none --import-source on -k regex:patch -s 2 -c 2 -o gpurun_out/sub3 python tools/profile_subband.py 3 > /dev/null 2>&1; ncu -i gpurun_out/sub3.ncu-rep --page raw --csv > gpurun_out/sub3_raw.csv
ncu --set full --clock-control none --import-source on -k regex:fir_tile -s 6 -c 3 -o gpurun_out/fir2 python tools/profile_fir.py > /dev/null 2>&1; ncu -i gpurun_out/fir2.ncu-rep --page raw --csv > gpurun_out/fir2_raw.csv
WICCA_ROWS_SUBBAND_DEPTHS=1,3,6 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rows2.csv python tools/bench_rows.py subbands wavelets > /dev/null 2>&1
ls -la gpurun_out | tail -8
