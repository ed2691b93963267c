# Round 2, GPU job 5 (1 GPU): K1 with adaptive chunks / 12 blocks per SM / auto key, the reworked bench line at N=1.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_trace.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -8 > gpurun_out/r02f_pytest.log; tail -8 gpurun_out/r02f_pytest.log
timeout 600 python tools/r02_tune.py --skip-k4 > gpurun_out/r02f_tune.log 2>&1; grep -v "^{" gpurun_out/r02f_tune.log | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; tail -3 gpurun_out/r02f_bench_n1.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02f_bench_n1.json').read().strip().splitlines()[-1])
    for k in ('metric','value','ms_per_iter','e2e','roofline','cpu_baseline','parity_checked','clocks','gpu_launches'): print(k, json.dumps(d.get(k))[:600])
    r=d['rays']
    for k in ('value','unordered_kernel','e2e','parity_checked','cpu_baseline','roofline'): print('rays.'+k, json.dumps(r.get(k))[:600])
    print('large', json.dumps(d.get('large_scene'))[:900])
except Exception as e: print('parse failed', e)
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-900
