# Round 2, 8-GPU job after the packed transfer streams: probe at 8 and 4 ranks, bench at N=8 (full) and N=4 / N=2 (short), in-process handle, group tests.
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 tools/multi_bounce_probe.py > gpurun_out/r02x_probe_n8.log 2>&1; grep "^192\|^256\|hierarchy" gpurun_out/r02x_probe_n8.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29725 tools/multi_bounce_probe.py > gpurun_out/r02x_probe_n4.log 2>&1; grep "^192\|^256\|hierarchy" gpurun_out/r02x_probe_n4.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02x_bench_n8.json 2> gpurun_out/r02x_bench_n8.err; tail -3 gpurun_out/r02x_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 4 --steps 10 --warmup 3 --no-large > gpurun_out/r02x_bench_n4.json 2> gpurun_out/r02x_bench_n4.err; tail -3 gpurun_out/r02x_bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29724 bench.py --gpus 2 --steps 10 --warmup 3 --no-large > gpurun_out/r02x_bench_n2.json 2> gpurun_out/r02x_bench_n2.err; tail -3 gpurun_out/r02x_bench_n2.err
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r02x_bench_n{n}.json').read().strip().splitlines() if l.startswith('{')][-1])
        print(n, 'value', d['value'], 'us/iter', d['ms_per_iter']*1e3, 'frac', d['roofline']['frac'], 'moved_frac', d['roofline']['moved_frac'], 'e2e', d['e2e']['value'], 'parity', json.dumps(d['parity_checked'])[:300])
        print(n, 'rays', d['rays']['value'], 'e2e', d['rays']['e2e']['value'], d['rays']['parity_checked'])
        if d.get('large_scene'): print(n, 'large', json.dumps(d['large_scene'])[:700])
    except Exception as e: print(n, 'parse failed', e)
PY
timeout 600 python tools/group_probe.py 8 > gpurun_out/r02x_group_n8.log 2>&1; tail -2 gpurun_out/r02x_group_n8.log | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -q 2>&1 | tail -4
