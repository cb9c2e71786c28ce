#!/bin/bash
N=$1; TAG=$2; PORT=29700
run() { name=$1; shift; PORT=$((PORT+1));
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --sustain-seconds 0 "$@" 2> gpurun_out/${TAG}_${name}.err | tail -1 > gpurun_out/${TAG}_${name}.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}.json").read())
    print("${name} N=$N: ms %.3f  G msg/s %.3f  checksum %s launch %s" % (d["ms_per_step"], d["value"]/1e9, d["checksum"], d["config"]["launch"][:30]))
except Exception as e:
    print("${name}: FAILED", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-800:])
PY
}
run peer_a --exchange peer
run peer_b --exchange peer
run peer_nograph --exchange peer --no-graph
run nccl --exchange nccl
run halo --exchange halo
