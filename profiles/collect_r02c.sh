#!/bin/bash
# Closing line of round 2 (after the C5 step diet and the narrow-row encode): tests, both bench arms, C5 launch list.
#   gpurun --timeout 900 -- 'bash profiles/collect_r02c.sh'
O=gpurun_out
T=r02
B="--no-cpu-baseline --no-e2e --cuda-graph 0 --extra-workloads= --steps 2 --warmup 3"
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${T}_pytest.log
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_launches_c5.csv \
    python bench.py --workload c5 --mode train $B > /dev/null 2>&1
tail -2 $O/${T}_pytest.log; head -c 300 $O/${T}_bench.json; echo; head -c 200 $O/${T}_bench_reference.json; echo; tail -2 $O/${T}_bench.err
