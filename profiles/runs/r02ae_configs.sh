# round 2, run ae: C1..C5 at full size on one GPU through the public API, final tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python profiles/run_configs.py > gpurun_out/r02ae_configs.json 2> gpurun_out/r02ae_configs.err
cat gpurun_out/r02ae_configs.json | cut -c1-330; tail -3 gpurun_out/r02ae_configs.err
