for p in 1 2 3; do for s in 3 5 9; do
python bench.py --steps 3 --warmup 3 --no-cpu --no-also --no-checks --e2e-pipes $p --e2e-slots $s > gpurun_out/e2e_${p}_${s}.json 2> gpurun_out/e2e_${p}_${s}.err
python -c "
import json; d=json.loads(open('gpurun_out/e2e_${p}_${s}.json').read().strip().splitlines()[-1]); print('pipes $p slots $s e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2))"
done; done
