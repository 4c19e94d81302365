set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --quick > gpurun_out/r2_bench_tma.json 2> gpurun_out/r2_bench_tma.err; tail -3 gpurun_out/r2_bench_tma.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_tma.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'].get('first_conv_fwd'), d['kernel_ms_total_per_step'])
i=d.get('inference',{})
print(i.get('value'), i.get('ms_per_call'), i.get('ms_per_call_repacking'), i.get('kernel_ms_per_call'))
PY
