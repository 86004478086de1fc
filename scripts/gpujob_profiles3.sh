# GPU box: --set full capture of every kernel of one final-configuration cfg4 sweep (after the command has exited 0 without ncu)
set -x
RSAC_GRAPH=0 python scripts/prof_run.py 1024 2 > gpurun_out/r2u_prof_plain.log 2>&1; tail -1 gpurun_out/r2u_prof_plain.log
RSAC_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"epnp_minimal_subwarp|score_kernel|ransac_select|select_eigen|early_exit_flag" -c 19 -o /tmp/r2u_sweep python scripts/prof_run.py 1024 1 > gpurun_out/r2u_ncu_sweep.log 2>&1
python scripts/ncu_summary.py /tmp/r2u_sweep.ncu-rep > gpurun_out/r2u_sweep_kernels_ncu_full.txt 2>&1
tail -3 gpurun_out/r2u_ncu_sweep.log
