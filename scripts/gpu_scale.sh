#!/bin/bash
# multi-GPU bench (strong scaling), launched like the driver does: $1 = ranks; writes gpurun_out/r02_scale_g$N.json
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_scale_g$N.json 2> gpurun_out/r02_scale_g$N.err
echo "rc=$?"; tail -n 3 gpurun_out/r02_scale_g$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_scale_g$N.json").read().strip().splitlines()[-1])
print(d["config"]["workload"], d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d.get("stages"))
for k, v in (d.get("workloads") or {}).items():
    print(k, "value", round(v["value"]), "ms/step", round(v["ms_per_step"], 3), "e2e", round(v["e2e"]["value"]))
PY
