#!/bin/bash
# Dev aid (GPU box): plain run, launch list, then one `ncu --set full` capture of the three bulk kernels of a 17-frame batch.
# Outputs under gpurun_out/: r02_plain2.log, r02_ncu_launches2.csv, r02_prof_bulk2.ncu-rep
CMD="python bench.py --lights 17 --ncal 4 --steps 1 --warmup 3 --no-cpu --no-e2e --no-also --no-checks --pipes 1"
$CMD > gpurun_out/r02_plain2.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_ncu_launches2.csv $CMD > gpurun_out/r02_ncu_l2.log 2>&1
echo "launch list rc=$?"
# matched launches per step: sky(reference), calibrate_batch, sky(batch), warp: skip step 0 and the next reference frame
ncu --set full --clock-control none --import-source on -k regex:"k_sky_sub_stats|k_calibrate_gray_ring|k_calibrate_gray_batch|k_warp_accumulate_tma" -s 5 -c 3 \
    -o gpurun_out/r02_prof_bulk2 -f $CMD > gpurun_out/r02_ncu_f2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/r02_prof_bulk2.ncu-rep gpurun_out/r02_ncu_launches2.csv
