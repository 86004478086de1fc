set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
python bench.py --impl reference > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; tail -c 300 gpurun_out/r2e_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/r2e_ncu_bench.log 2>&1
RSAC_PROF_NEW=1 ncu --set full --clock-control none --import-source on -k regex:"kfdb|sim3_search|proj_|mlpnp_minimal_range|pack_pnp_indexed|pnp_keypoint" -c 40 -o /tmp/r2e_new python scripts/prof_run.py 64 1 > gpurun_out/r2e_ncu_new.log 2>&1
python scripts/ncu_summary.py /tmp/r2e_new.ncu-rep > gpurun_out/r2e_new_ncu_full.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ransac_select|epnp_minimal_subwarp|score_kernel" -c 12 -o /tmp/r2e_sweep python scripts/prof_run.py 1024 1 > gpurun_out/r2e_ncu_sweep.log 2>&1
python scripts/ncu_summary.py /tmp/r2e_sweep.ncu-rep > gpurun_out/r2e_sweep_ncu_full.txt 2>&1
ls -la gpurun_out | tail -12
