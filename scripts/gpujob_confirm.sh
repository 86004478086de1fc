set -x
cd $GRAFT_REPO_ROOT
( time python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" ) > gpurun_out/r2ad_smoke.log 2>&1
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2ad_bench_n1_reference.json 2> gpurun_out/r2ad_bench_ref.err
( time python bench.py ) > gpurun_out/r2ad_bench_n1.json 2> gpurun_out/r2ad_bench.err
tail -3 gpurun_out/r2ad_smoke.log; tail -4 gpurun_out/r2ad_bench_ref.err; tail -4 gpurun_out/r2ad_bench.err
cut -c1-300 gpurun_out/r2ad_bench_n1.json
