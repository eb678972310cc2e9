#!/bin/bash
# Dev aid: device-resident config-2 throughput for several SM partitions (OSS_B200_SPLIT = SMs of the sky kernel's green context).
for s in ${SPLITS:-0 80 88 96 104 112}; do
  echo "== OSS_B200_SPLIT=$s"
  OSS_B200_SPLIT=$s python bench.py --no-cpu --no-e2e --no-also --no-checks --steps ${STEPS:-5} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        b = json.loads(l); k = b['roofline']['kernels']
        print('value %.1f fps  ms/step %.3f | serial us/frame:' % (b['value'], b['ms_per_step']), {n: round(v['us_per_frame'], 1) for n, v in k.items()})
    elif l: print(l)
"
done
