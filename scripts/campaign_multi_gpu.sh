#!/bin/bash
# multi-GPU campaign (gpurun --gpus N -- bash scripts/campaign_multi_gpu.sh N): two-rank tests (N == 2) and the bench line with its parity pre-check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
if [ "$N" = "2" ]; then ( timeout 900 python -m pytest tests/test_gpu_variants.py -q -m gpu -k two_ranks 2>&1 | tail -3 ); fi
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/m_bench$N.json 2> gpurun_out/m_bench$N.err
tail -5 gpurun_out/m_bench$N.err; python - <<PY
import json
txt=open("gpurun_out/m_bench$N.json").read()
line=[l for l in txt.splitlines() if l.startswith("{")][-1]
d=json.loads(line)
print("ms/step", d["ms_per_step"], "value", d["value"])
print("parity", json.dumps(d["parity"])[:600])
print("kernel_ms", json.dumps(d["kernel_ms_per_step"]))
print("nvlink", json.dumps(d.get("nvlink"))[:300])
print("secondary", json.dumps(d.get("secondary"))[:500])
PY
