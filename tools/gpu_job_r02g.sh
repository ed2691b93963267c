# Round 2, GPU job 6 (2 GPUs): the multi-GPU gather on real NVLink -- parity test, switches, bench at N=2.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15 > gpurun_out/r02g_pytest_multi.log; tail -15 gpurun_out/r02g_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/multi_bounce_probe.py > gpurun_out/r02g_probe_n2.log 2>&1; tail -12 gpurun_out/r02g_probe_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 10 --warmup 3 --no-large > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; tail -5 gpurun_out/r02g_bench_n2.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02g_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
    for k in ('metric','value','ms_per_iter','e2e','roofline','parity_checked','row_blocks','gpu_launches'): print(k, json.dumps(d.get(k))[:700])
    r=d['rays']
    for k in ('value','e2e','parity_checked'): print('rays.'+k, json.dumps(r.get(k))[:500])
except Exception as e: print('parse failed', e)
PY
