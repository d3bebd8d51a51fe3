#!/bin/bash
# Round-2 closing artefacts after the tc5_eval hand-off fix / link arithmetic / narrow-row plan (subset of
# collect_r02.sh; run from the repo root under gpurun, everything lands in gpurun_out/):
#   gpurun --timeout 1500 -- 'bash profiles/collect_r02b.sh'
O=gpurun_out
T=r02
B="--no-cpu-baseline --no-e2e --cuda-graph 0 --extra-workloads= --steps 2 --warmup 3"
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${T}_pytest.log
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
for w in c3 c5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_launches_$w.csv \
      python bench.py --workload $w --mode train $B > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc5_eval_kernel -s 3 -c 1 -f -o $O/${T}_tc5_eval_c3 \
    python bench.py --workload c3 --mode eval $B > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:tc5_encode_kernel|tc5_encode_bwd_kernel|link_stream_kernel' -s 9 -c 3 -f -o $O/${T}_c3_train_kernels \
    python bench.py --workload c3 --mode train $B > /dev/null 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_r02.py rest \
    > $O/${T}_memcheck_rest.log 2>&1; echo "memcheck exit $?" >> $O/${T}_memcheck_rest.log
tail -3 $O/${T}_pytest.log; head -c 1200 $O/${T}_bench.json; echo; head -c 500 $O/${T}_bench_reference.json; echo
tail -n 4 $O/${T}_memcheck_rest.log
ls -la $O | tail -15
