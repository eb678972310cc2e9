#!/bin/bash
# Dev aid (GPU box): A/B builds of liboss_b200.so under build/variants/liboss_<name>.so — parity tests of the sky kernel, then
# per-kernel times from a short bench run.  Usage: tools/variant_sweep.sh A B C ...   (results: gpurun_out/var_<name>.*)
LIB=openskystacker_b200/csrc/liboss_b200.so
cp $LIB /tmp/liboss_keep.so
for v in "$@"; do
  cp build/variants/liboss_$v.so $LIB
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${VARIANT_TESTS:-sky_subtraction or star_lists_identical}" > gpurun_out/var_$v.test 2>&1
  echo "variant $v tests rc=$? $(tail -1 gpurun_out/var_$v.test)"
  timeout 600 python bench.py --steps ${VARIANT_STEPS:-5} --warmup 3 --no-cpu --no-e2e --no-also --no-checks ${VARIANT_BENCH_ARGS} > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/var_$v.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print("variant $v value %.1f fps  ms/step %.3f  " % (d["value"], d["ms_per_step"]) + "  ".join("%s %.1f us" % (n, k[n]["us_per_frame"]) for n in k), " serial %.2f ms" % d["roofline"]["serial_step_ms"])
except Exception as e:
    print("variant $v bench failed", e)
PY
done
cp /tmp/liboss_keep.so $LIB
