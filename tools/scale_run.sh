#!/bin/bash
# multi-GPU bench sweep on one box: usage tools/scale_run.sh N tag ; writes gpurun_out/<tag>_<name>.json
N=$1; TAG=$2; PORT=29600
run() { name=$1; shift; PORT=$((PORT+1));
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --sustain-seconds 0 "$@" 2> gpurun_out/${TAG}_${name}.err | tail -1 > gpurun_out/${TAG}_${name}.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}.json").read())
    print("${name} N=$N: ms %.3f  G msg/s %.3f  frac/gpu %.3f  checksum %s  %s" % (d["ms_per_step"], d["value"]/1e9, d["roofline"]["frac"], d["checksum"], json.dumps(d.get("exchange",""))[:200]))
except Exception as e:
    print("${name}: FAILED", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-800:])
PY
}
run cfg4_band --config cfg4 --local-band 512
run cfg4_rand --config cfg4
run cfg2_rand_halo --exchange halo
run cfg2_rand_peer --exchange peer
run cfg2_band --local-band 512
run cfg3 --config cfg3
