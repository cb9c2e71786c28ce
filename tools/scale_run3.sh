#!/bin/bash
PORT=29800
one() { timeout 300 python bench.py --steps 5 --warmup 3 --sustain-seconds 0 --no-cpu-baseline --no-e2e --no-aten-baseline "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', sys.argv[1:], 'checksum', d['checksum'], 'ms', round(d['ms_per_step'],3))" "$@"; }
multi() { N=$1; shift; PORT=$((PORT+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 5 --warmup 3 --sustain-seconds 0 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$N', sys.argv[1:], 'checksum', d['checksum'], 'ms', round(d['ms_per_step'],3))" "$@"; }
one
multi 2 --exchange peer
multi 2 --exchange halo
one --config cfg4 --vars 100000 --pairwise 300000 --high 50000 --local-band 512
multi 2 --config cfg4 --vars 100000 --pairwise 300000 --high 50000 --local-band 512
