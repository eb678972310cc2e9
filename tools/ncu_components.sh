#!/bin/bash
# Dev aid (GPU box): one `ncu --set full` capture of a 10-frame k_components launch of config 4 (61 MP mono, threshold 5: thousands of
# components per frame) -> gpurun_out/<name>.ncu-rep
NAME=${1:-r02_prof_components}
CMD="python bench.py --config 4 --lights 10 --steps 1 --warmup 3 --no-cpu --no-e2e --no-also --no-checks --pipes 1"
$CMD > gpurun_out/${NAME}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${NAME}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"k_components$" -s 6 -c 2 -o gpurun_out/$NAME -f $CMD > gpurun_out/${NAME}_ncu.log 2>&1
echo "capture rc=$?"; ls -la gpurun_out/$NAME.ncu-rep
