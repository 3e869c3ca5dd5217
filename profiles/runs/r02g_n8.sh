set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NGPU:-8}
python - <<'PY' > gpurun_out/r02g_multicast_probe.txt 2>&1
import torch
from torch.distributed._symmetric_memory import _SymmetricMemory as S
try:
    print('has_multicast_support', S.has_multicast_support(torch.device('cuda').type, 0))
except Exception as e:
    print('probe error', repr(e))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/dist_check.py > gpurun_out/r02g_dist_check_n$N.json 2> gpurun_out/r02g_dist_check_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02g_bench_n$N.json 2> gpurun_out/r02g_bench_n$N.err
RAIDER_B200_NO_MULTICAST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02g_bench_n${N}_nomc.json 2> gpurun_out/r02g_bench_n${N}_nomc.err
RAIDER_B200_NO_MULTICAST=1 RDR_K3_TILE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02g_bench_n${N}_nomc_tile0.json 2> gpurun_out/r02g_bench_n${N}_nomc_tile0.err
cat gpurun_out/r02g_multicast_probe.txt; cat gpurun_out/r02g_dist_check_n$N.json | cut -c1-400; tail -3 gpurun_out/r02g_bench_n$N.err
