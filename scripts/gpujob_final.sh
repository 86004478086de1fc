# GPU box: the round's verification pass -- GPU suite (default and checked builds), smoke, both bench arms, launch list
set -x
P=${1:-r2s}
python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; tail -3 gpurun_out/${P}_pytest.log
bash scripts/run_checked.sh > gpurun_out/${P}_pytest_checked.log 2>&1; tail -3 gpurun_out/${P}_pytest_checked.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${P}_smoke.log 2>&1; tail -2 gpurun_out/${P}_smoke.log
python bench.py --impl reference > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err
python bench.py > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err; tail -c 300 gpurun_out/${P}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/${P}_ncu_bench.log 2>&1
