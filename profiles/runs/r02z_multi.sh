# round 2, run z: multi-GPU lines from the final tree.  usage: NGPU=2|4|8 bash profiles/runs/r02z_multi.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NGPU:-8}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 ${@:3} > gpurun_out/r02z_$2_n$N.json 2> gpurun_out/r02z_$2_n$N.err; }
run 29511 dist_check profiles/dist_check.py
run 29512 bench bench.py --gpus $N --steps 10 --warmup 3
run 29513 bench_c5 bench.py --gpus $N --steps 5 --warmup 3 --config c5
if [ "$N" = "8" ]; then
run 29514 bench_c5ml145 bench.py --gpus $N --steps 5 --warmup 3 --config c5-ml145
run 29516 bench_run2 bench.py --gpus $N --steps 10 --warmup 3
fi
cut -c1-300 gpurun_out/r02z_dist_check_n$N.json; tail -2 gpurun_out/r02z_bench_n$N.err; tail -2 gpurun_out/r02z_bench_c5_n$N.err
