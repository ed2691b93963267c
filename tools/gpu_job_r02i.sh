# Round 2, GPU job 8 (8 GPUs): the gather at 8 ranks on NVSwitch -- probe of the switches, the bench line at N=8, the in-process handle.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/r02i_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/multi_bounce_probe.py > gpurun_out/r02i_probe_n8.log 2>&1; grep -v "^\*\|OMP_NUM" gpurun_out/r02i_probe_n8.log | tail -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02i_bench_n8.json 2> gpurun_out/r02i_bench_n8.err; tail -5 gpurun_out/r02i_bench_n8.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02i_bench_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
    for k in ('metric','value','ms_per_iter','e2e','roofline','parity_checked','row_blocks','gpu_launches','clocks'): print(k, json.dumps(d.get(k))[:800])
    r=d['rays']
    for k in ('value','e2e','parity_checked'): print('rays.'+k, json.dumps(r.get(k))[:500])
    print('large', json.dumps(d.get('large_scene'))[:900])
except Exception as e: print('parse failed', e)
PY
timeout 600 python tools/group_probe.py 8 > gpurun_out/r02i_group_n8.log 2>&1; tail -3 gpurun_out/r02i_group_n8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29713 tools/multi_bounce_probe.py > gpurun_out/r02i_probe_n4.log 2>&1; grep -v "^\*\|OMP_NUM" gpurun_out/r02i_probe_n4.log | tail -9
