#!/bin/bash
# Produces the round-1 measurement artefacts on a B200 box (run from the repo root under
# gpurun; everything lands in gpurun_out/, summaries are then written to profiles/ with
# profiles/summarize.py in the build container):
#   gpurun --timeout 1500 -- 'bash profiles/collect_r01.sh'
O=gpurun_out
T=r01
B="--no-cpu-baseline --no-e2e --cuda-graph 0 --extra-workloads= --steps 2 --warmup 3"
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${T}_pytest.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
# launch lists (kernel shares of a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_launches_c4.csv \
    python bench.py $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_launches_c3.csv \
    python bench.py --workload c3 --mode train $B > /dev/null 2>&1
# full captures of the dominant kernels (gpurun_out/ must stay below 64 MiB: no source import for
# the multi-kernel C3 report)
ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 6 -c 1 -f -o $O/${T}_fused2_eval_c4 \
    python bench.py --mode eval $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 6 -c 1 -f -o $O/${T}_fused2_train_c4 \
    python bench.py --mode train $B > /dev/null 2>&1
ncu --set full --clock-control none -k 'regex:_mma_kernel|link_stream' -s 3 -c 3 -f -o $O/${T}_c3_kernels \
    python bench.py --workload c3 --mode train $B > /dev/null 2>&1
# the multi-kernel report embeds ~60 MB of SASS: keep its raw metric page only
ncu -i $O/${T}_c3_kernels.ncu-rep --page raw --csv > $O/${T}_c3_kernels_raw.csv 2>/dev/null && rm -f $O/${T}_c3_kernels.ncu-rep
# memory / race checkers on the new kernel families (small parity cases)
K='vs_oracle and (256-333-5 or 128-77-8 or 1000-95-1 or 200-150-2) and (composed or slab)'
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" \
    > $O/${T}_memcheck.log 2>&1; echo "memcheck exit $?" >> $O/${T}_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" \
    > $O/${T}_racecheck.log 2>&1; echo "racecheck exit $?" >> $O/${T}_racecheck.log
tail -3 $O/${T}_pytest.log; cat $O/${T}_bench.json; cat $O/${T}_bench_reference.json; for f in $O/${T}_memcheck.log $O/${T}_racecheck.log; do echo "== $f"; tail -n 6 $f; done
ls -la $O | tail -15
