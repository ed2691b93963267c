set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02o_pytest_gpu.log; tail -6 gpurun_out/r02o_pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err ) 2>&1 | tail -3; tail -3 gpurun_out/r02o_bench_n1.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02o_bench_n1.json').read().strip().splitlines() if l.startswith('{')][-1])
    for k in ('metric','value','ms_per_iter','e2e','roofline','parity_checked','gpu_launches','clocks'): print(k, json.dumps(d.get(k))[:500])
    r=d['rays']
    for k in ('value','unordered_kernel','e2e','parity_checked'): print('rays.'+k, json.dumps(r.get(k))[:500])
    print('large', json.dumps(d.get('large_scene'))[:700])
    print('kd', json.dumps(d['bsp_side'].get('kd_build_fast',{}).get('C1_box_room'))[:900])
except Exception as e: print('parse failed', e)
PY
( time timeout 300 python bench.py --impl reference --steps 10 --warmup 3 ) 2>&1 | cut -c1-300 | tail -5
