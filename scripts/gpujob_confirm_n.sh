# both arms at N GPUs (argument), launched the way the driver launches them
N=${1:-2}
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
( time timeout 200 $TR bench.py --impl reference --gpus $N --steps 5 --warmup 1 ) > gpurun_out/r2aj_bench_n${N}_reference.json 2> gpurun_out/r2aj_ref_n${N}.err
( time timeout 300 $TR bench.py --gpus $N ) > gpurun_out/r2aj_bench_n${N}.json 2> gpurun_out/r2aj_bench_n${N}.err
tail -4 gpurun_out/r2aj_ref_n${N}.err; tail -4 gpurun_out/r2aj_bench_n${N}.err
cut -c1-300 gpurun_out/r2aj_bench_n${N}.json; cut -c1-200 gpurun_out/r2aj_bench_n${N}_reference.json
