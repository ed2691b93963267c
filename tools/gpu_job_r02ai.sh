set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29724 bench.py --gpus 2 --no-large --no-cpu --steps 5 > gpurun_out/r02ai_bench_n2.json 2> gpurun_out/r02ai_bench_n2.err; tail -2 gpurun_out/r02ai_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02ai_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'us/iter', d['ms_per_iter']*1e3, d['roofline']['kernel'], 'parity', json.dumps(d['parity_checked'])[:330])
PY
