# GPU box: the round's verification pass -- GPU suite (default and checked builds), smoke, both bench arms
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
bash scripts/run_checked.sh > gpurun_out/r2m_pytest_checked.log 2>&1; tail -3 gpurun_out/r2m_pytest_checked.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
python bench.py --impl reference > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err
python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; tail -c 300 gpurun_out/r2m_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/r2m_ncu_bench.log 2>&1
