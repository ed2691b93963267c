# Round 2, 4-GPU job: the block-row gather with real peers at 4 ranks, the packed multi-GPU path forced on the same box, multi-GPU tests.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 4 --no-large > gpurun_out/r02af_bench_n4.json 2> gpurun_out/r02af_bench_n4.err; tail -2 gpurun_out/r02af_bench_n4.err
VRAD_K4_PACK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29726 bench.py --gpus 4 --no-large --steps 5 > gpurun_out/r02af_bench_n4_packed.json 2> gpurun_out/r02af_bench_n4_packed.err; tail -2 gpurun_out/r02af_bench_n4_packed.err
python - <<'PY'
import json
for f in ('r02af_bench_n4', 'r02af_bench_n4_packed'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d['value'], 'us/iter', d['ms_per_iter']*1e3, 'frac', d['roofline']['frac'], 'moved', d['roofline']['moved_frac'], d['roofline']['kernel'], 'e2e', d['e2e']['value'], 'parity', json.dumps(d['parity_checked'])[:330], d['row_blocks'])
    except Exception as e: print(f, 'parse failed', e)
PY
