#!/bin/bash
# Final round-2 evidence on one B200 (outputs under gpurun_out/, copied to profiles/ by hand):  bash tools/r2_final.sh
set -u
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err ) 2> gpurun_out/r2f_bench_n1.time; echo "bench rc=$?"; tail -n 3 gpurun_out/r2f_bench_n1.time
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err ) 2> gpurun_out/r2f_bench_ref.time; echo "reference arm rc=$?"; tail -n 3 gpurun_out/r2f_bench_ref.time
python tools/bench_epilogue.py > gpurun_out/r2f_epilogue.jsonl 2> gpurun_out/r2f_epilogue.err; echo "epilogue rc=$?"
ncu --set full --clock-control none --import-source on -k regex:resize_area_rows -s 1 -c 4 -o gpurun_out/r2f_epi -f python tools/profile_epilogue.py > /dev/null 2>&1
ncu -i gpurun_out/r2f_epi.ncu-rep --page raw --csv > gpurun_out/r2f_epilogue_rows_full_raw.csv
WICCA_ROWS_SUBBAND_DEPTHS=1,3,6 python tools/bench_rows.py > gpurun_out/r2f_rows.jsonl 2> gpurun_out/r2f_rows.err; echo "rows rc=$?"; tail -n 3 gpurun_out/r2f_rows.err
rm -f gpurun_out/r2f_epi.ncu-rep
ls -la gpurun_out | grep r2f_
