# both arms at N = 2, launched the way the driver launches them
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
( time $TR bench.py --impl reference --gpus 2 --steps 5 --warmup 1 ) > gpurun_out/r2ai_bench_n2_reference.json 2> gpurun_out/r2ai_ref.err
( time $TR bench.py --gpus 2 ) > gpurun_out/r2ai_bench_n2.json 2> gpurun_out/r2ai_bench.err
tail -4 gpurun_out/r2ai_ref.err; tail -4 gpurun_out/r2ai_bench.err
cut -c1-400 gpurun_out/r2ai_bench_n2.json; cut -c1-200 gpurun_out/r2ai_bench_n2_reference.json
