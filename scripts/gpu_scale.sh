#!/bin/bash
# multi-GPU bench (strong scaling): $1 = ranks, $2 = workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$1; W=${2:-dhe1m}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --workload $W > gpurun_out/scale_${W}_g$N.json 2> gpurun_out/scale_${W}_g$N.err
echo "rc=$?"; tail -n 3 gpurun_out/scale_${W}_g$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/scale_${W}_g$N.json").read().strip().splitlines()[-1])
print("$W", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d.get("stages"))
PY
