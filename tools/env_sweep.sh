#!/bin/bash
# Dev aid (GPU box): the short bench under several environment settings.  Usage: tools/env_sweep.sh "A=1 B=0" "A=0" ...
i=0
for e in "$@"; do
  i=$((i+1))
  env $e timeout 600 python bench.py --steps ${SWEEP_STEPS:-5} --warmup 3 --no-cpu --no-e2e --no-also --no-checks ${SWEEP_ARGS} > gpurun_out/env_$i.json 2> gpurun_out/env_$i.err
  python - "$e" gpurun_out/env_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print("[%s] value %.1f fps  ms/step %.3f  " % (sys.argv[1], d["value"], d["ms_per_step"]) + "  ".join("%s %.1f us" % (n, k[n]["us_per_frame"]) for n in k), " serial %.2f ms" % d["roofline"]["serial_step_ms"])
except Exception as ex:
    print("[%s] bench failed" % sys.argv[1], ex); print(open(sys.argv[2].replace(".json", ".err")).read()[-800:])
PY
done
