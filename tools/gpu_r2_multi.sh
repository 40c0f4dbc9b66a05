#!/bin/bash
# Multi-GPU call: bench.py under torchrun at N = $1 ranks (full path + matcher regions), then run_test.py dataset mode at N ranks.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
nproc > gpurun_out/r02_multi_n${N}_host.txt; nvidia-smi topo -m >> gpurun_out/r02_multi_n${N}_host.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/r02_bench_n${N}.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/gpu_dataset_mode.py --repeat 43 --workers 4 > gpurun_out/r02_dataset_n${N}.json 2> gpurun_out/r02_dataset_n${N}.err; echo "dataset N=$N exit $?"; cat gpurun_out/r02_dataset_n${N}.json
python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r02_bench_n${N}.json").read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "n_gpus", "ms_per_step", "status", "clocks", "affinity", "h2d_probe")})
    print("e2e", l["e2e"])
    for m in ("matcher_config2", "matcher_config5"):
        if l.get(m): print(m, {k: l[m].get(k) for k in ("value", "ms_per_step", "region_s", "e2e")})
    for r in ("roofline", "roofline_config5"):
        if l.get(r): print(r, {k: l[r][k] for k in ("achieved", "peak", "frac", "frac_of_burst_peak")})
except Exception as e:
    print("unreadable", e)
PY
