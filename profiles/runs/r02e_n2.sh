set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02e_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/dist_check.py > gpurun_out/r02e_dist_check_n2.json 2> gpurun_out/r02e_dist_check_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err
RAIDER_B200_T_BUDGET_GB=100 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --config c5 > gpurun_out/r02e_bench_c5_n2.json 2> gpurun_out/r02e_bench_c5_n2.err
cat gpurun_out/r02e_dist_check_n2.json; tail -5 gpurun_out/r02e_dist_check_n2.err; tail -3 gpurun_out/r02e_bench_n2.err; tail -3 gpurun_out/r02e_bench_c5_n2.err
