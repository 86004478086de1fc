# GPU box: final-configuration captures of the sweep's kernels (after the commands have exited 0 without ncu)
set -x
python scripts/prof_run.py 1024 2 > gpurun_out/r2h_prof_plain.log 2>&1; tail -1 gpurun_out/r2h_prof_plain.log
ncu --set full --clock-control none --import-source on -k regex:"epnp_minimal_subwarp" -c 5 -o /tmp/r2h_solve python scripts/prof_run.py 1024 1 > gpurun_out/r2h_ncu_solve.log 2>&1
python scripts/ncu_summary.py /tmp/r2h_solve.ncu-rep > gpurun_out/r2h_solve_ncu_full.txt 2>&1
ncu -i /tmp/r2h_solve.ncu-rep --page raw --csv > gpurun_out/r2h_solve_raw.csv 2>/dev/null
python bench.py --impl reference > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -c 300 gpurun_out/r2h_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/r2h_ncu_bench.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
