#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): bench records, ncu launch list, per-kernel DRAM traffic, full captures.
set -u
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r02_bench_cfg2.json 2> $O/r02_bench_cfg2.err
python bench.py --config cfg3 --steps 20 --warmup 5 > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err
python bench.py --edge-types 4 --steps 20 --warmup 5 --no-cpu-baseline --no-aten-baseline > $O/r02_bench_cfg2_T4.json 2> /dev/null
python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-aten-baseline --no-e2e > $O/r02_bench_cfg5.json 2> /dev/null
python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-aten-baseline > $O/r02_bench_cfg4_1gpu_random.json 2> /dev/null
python bench.py --config cfg4 --local-band 512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-aten-baseline > $O/r02_bench_cfg4_1gpu_band512.json 2> /dev/null
# launch list + DRAM traffic of one step (no graph: every launch is a plain kernel node)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mp_|w_split|et_permute' --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-aten-baseline --sustain-seconds 0 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'mp_' -s 100 -c 28 --csv --log-file $O/r02_traffic.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-aten-baseline --sustain-seconds 0 > /dev/null 2>&1
# full captures: one layer of cfg2 (the 4 calls = 7 launches), one cfg3 fused call, one cfg3 destination-stationary call
ncu --set full --clock-control none --import-source on -k regex:'mp_' -s 100 -c 7 -o $O/r02_full_cfg2 -f \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-aten-baseline --sustain-seconds 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mp_' -s 120 -c 2 -o $O/r02_full_cfg3_fused -f \
    python bench.py --config cfg3 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-aten-baseline --sustain-seconds 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mp_' -s 120 -c 2 -o $O/r02_full_cfg3_dst -f \
    python bench.py --config cfg3 --src-calls none --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-aten-baseline --sustain-seconds 0 > /dev/null 2>&1
ls -la $O | tail -20
