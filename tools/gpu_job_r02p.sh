# Round 2, final 8-GPU job: probe with NVLink byte counters around it, bench at N=8 (full) and N=4 / N=2 (short), the in-process handle, group tests.
set -x
mkdir -p gpurun_out
python tools/nvlink_counters.py > gpurun_out/r02p_nvlink_before.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 tools/multi_bounce_probe.py > gpurun_out/r02p_probe_n8.log 2>&1; grep -v "^\*\|OMP_NUM" gpurun_out/r02p_probe_n8.log | tail -12 | cut -c1-400
python tools/nvlink_counters.py > gpurun_out/r02p_nvlink_after.json; cat gpurun_out/r02p_nvlink_before.json | cut -c1-300; cat gpurun_out/r02p_nvlink_after.json | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err; tail -3 gpurun_out/r02p_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 4 --steps 10 --warmup 3 --no-large > gpurun_out/r02p_bench_n4.json 2> gpurun_out/r02p_bench_n4.err; tail -3 gpurun_out/r02p_bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29724 bench.py --gpus 2 --steps 10 --warmup 3 --no-large > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err; tail -3 gpurun_out/r02p_bench_n2.err
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r02p_bench_n{n}.json').read().strip().splitlines() if l.startswith('{')][-1])
        print(n, 'value', d['value'], 'us/iter', d['ms_per_iter']*1e3, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', json.dumps(d['parity_checked'])[:300])
        print(n, 'rays', d['rays']['value'], 'e2e', d['rays']['e2e']['value'], d['rays']['e2e']['coordinates']['value'], d['rays']['parity_checked'])
        if d.get('large_scene'): print(n, 'large', json.dumps(d['large_scene'])[:700])
    except Exception as e: print(n, 'parse failed', e)
PY
timeout 600 python tools/group_probe.py 8 > gpurun_out/r02p_group_n8.log 2>&1; tail -2 gpurun_out/r02p_group_n8.log | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -q 2>&1 | tail -4
