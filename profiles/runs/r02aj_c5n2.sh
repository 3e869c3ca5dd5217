# round 2, run aj: C5 at N = 2 once more (the e2e figure of r02ad_bench_c5_n2.json, 464 ms, was a slow host)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --config c5 > gpurun_out/r02aj_bench_c5_n2.json 2> gpurun_out/r02aj_bench_c5_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02aj_bench_c5_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
P
