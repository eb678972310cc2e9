#!/bin/bash
# dev sweep: frames in flight x pipelines, device-resident value only
for cfg in "10 3" "17 3" "25 2" "25 3" "17 2"; do
  set -- $cfg
  echo "slots=$1 pipes=$2" 
  OSS_B200_PIPES=$2 timeout 300 python bench.py --slots $1 --no-cpu --no-e2e --steps 3 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['serial_step_ms'], {k:round(v['us_per_frame'],1) for k,v in d['roofline']['kernels'].items()})"
done
