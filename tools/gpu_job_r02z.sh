set -x
mkdir -p gpurun_out
for P in 1 0 1 0; do
VRAD_K4_PACK=$P timeout 600 python bench.py --steps 5 --warmup 3 --no-large --no-cpu 2>/dev/null | python -c "
import sys, json
d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
print('PACK=$P', 'us/iter', d['ms_per_iter']*1e3, 'value', d['value'], 'moved_frac', d['roofline']['moved_frac'], d['clocks'])
"
done
timeout 300 python tools/k4_pack_probe.py 2>&1 | tail -11
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,clocks.max.mem,temperature.gpu --format=csv
