# Round 2, final 1-GPU job: the whole GPU suite, smoke, bench (both arms), and the launch list of the bench command.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --impl reference > gpurun_out/r02y_bench_ref.json 2> gpurun_out/r02y_bench_ref.err; tail -c 600 gpurun_out/r02y_bench_ref.json
timeout 900 python bench.py > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err; tail -3 gpurun_out/r02y_bench_n1.err; tail -c 300 gpurun_out/r02y_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-large --no-cpu > gpurun_out/r02y_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r02y_bench_under_ncu.log | cut -c1-300; wc -l gpurun_out/r02_final_launches.csv
python tools/ncu_summary.py gpurun_out/r02_final_launches.csv > gpurun_out/r02_final_launches_summary.txt 2>&1; head -20 gpurun_out/r02_final_launches_summary.txt
