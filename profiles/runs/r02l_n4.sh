set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/dist_check.py > gpurun_out/r02l_dist_check_n$N.json 2> gpurun_out/r02l_dist_check_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02l_bench_n$N.json 2> gpurun_out/r02l_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --config c5 > gpurun_out/r02l_bench_c5_n$N.json 2> gpurun_out/r02l_bench_c5_n$N.err
cut -c1-300 gpurun_out/r02l_dist_check_n$N.json; tail -2 gpurun_out/r02l_bench_n$N.err; tail -2 gpurun_out/r02l_bench_c5_n$N.err
