# Round 2, final 1-GPU job (block-row gather): the whole GPU suite, smoke, bench (both arms), the launch list of the bench command, ncu of the gather.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r02ah_bench_ref.json 2> gpurun_out/r02ah_bench_ref.err; tail -c 300 gpurun_out/r02ah_bench_ref.json
timeout 900 python bench.py > gpurun_out/r02ah_bench_n1.json 2> gpurun_out/r02ah_bench_n1.err; tail -3 gpurun_out/r02ah_bench_n1.err; tail -c 200 gpurun_out/r02ah_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-large --no-cpu > gpurun_out/r02ah_bench_under_ncu.log 2>&1; wc -l gpurun_out/r02_final_launches.csv
python tools/launch_summary.py gpurun_out/r02_final_launches.csv > gpurun_out/r02_final_launches_summary.txt 2>&1; head -8 gpurun_out/r02_final_launches_summary.txt
cat > /tmp/bk_target.py <<'PY'
import sys
sys.path.insert(0, '.')
import torch
from vrad_b200 import scenes
from vrad_b200.environment import environment_from_scene
s = scenes.multi_room()
env = environment_from_scene(s)
env.set_option("k4_graph", 0)
env.build_transfers(s.pvs)
e0 = torch.full((s.n_patches, 3), 100.0, device="cuda"); o = torch.empty_like(e0)
env.bounce(e0, 3, out=o, want_added=False)
env.close()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k4_gather_blocked' -c 3 -o gpurun_out/r02_k4_block_full python /tmp/bk_target.py > gpurun_out/r02ah_ncu.log 2>&1; tail -1 gpurun_out/r02ah_ncu.log
