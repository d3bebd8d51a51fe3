#!/bin/bash
# Round-2 measurement artefacts on a B200 box (run from the repo root under gpurun; everything lands
# in gpurun_out/, summaries are then written to profiles/ with profiles/summarize.py in the build
# container):
#   gpurun --timeout 2400 -- 'bash profiles/collect_r02.sh'
O=gpurun_out
T=r02
B="--no-cpu-baseline --no-e2e --cuda-graph 0 --extra-workloads= --steps 2 --warmup 3"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${T}_pytest.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
timeout 600 python profiles/kernel_bench.py > $O/${T}_kernel_bench.json 2>> $O/${T}_bench.err
timeout 600 python profiles/epoch_time.py 2>> $O/${T}_bench.err | grep '^{' > $O/${T}_epoch_time.jsonl
timeout 300 python profiles/host_pack_bench.py > $O/${T}_host_pack.log 2>> $O/${T}_bench.err
# launch lists (kernel shares of a step)
for w in c4 c2 c3 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_launches_$w.csv \
      python bench.py --workload $w --mode train $B > /dev/null 2>&1
done
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 6 -c 1 -f -o $O/${T}_fused2_eval_c4 \
    python bench.py --mode eval $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused3_kernel -s 6 -c 1 -f -o $O/${T}_fused3_train_c4 \
    python bench.py --mode train $B > /dev/null 2>&1
VIBO_DISABLE_FUSED3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 6 -c 1 -f -o $O/${T}_fused2_train_c4 \
    python bench.py --mode train $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc5_eval_kernel -s 3 -c 1 -f -o $O/${T}_tc5_eval_c3 \
    python bench.py --workload c3 --mode eval $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tc5_encode_kernel|tc5_encode_bwd_kernel|link_stream_kernel' -s 9 -c 3 -f -o $O/${T}_c3_train_kernels \
    python bench.py --workload c3 --mode train $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k 'regex:percell_mlp|sample_loop|step_tail|param_forward|adam_kernel|fused_finalize|unpack' -c 12 -f -o $O/${T}_new_kernels \
    python profiles/ncu_new_kernels.py > /dev/null 2>&1
ncu -i $O/${T}_new_kernels.ncu-rep --page raw --csv > $O/${T}_new_kernels_raw.csv 2>/dev/null && rm -f $O/${T}_new_kernels.ncu-rep
# memory / race checkers: the headline single-pass kernel with >= 3 ring laps per team, and the new kernels
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_r02.py \
    > $O/${T}_memcheck.log 2>&1; echo "memcheck exit $?" >> $O/${T}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_r02.py rest \
    > $O/${T}_racecheck_rest.log 2>&1; echo "racecheck exit $?" >> $O/${T}_racecheck_rest.log
VIBO_E5_DEBUG=8 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_r02.py rest \
    > $O/${T}_racecheck_rest_lockstep.log 2>&1; echo "racecheck exit $?" >> $O/${T}_racecheck_rest_lockstep.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_r02.py fused \
    > $O/${T}_racecheck_fused.log 2>&1; echo "racecheck exit $?" >> $O/${T}_racecheck_fused.log
VIBO_FUSED_DEBUG=3 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_r02.py fused \
    > $O/${T}_racecheck_fused_barrier.log 2>&1; echo "racecheck exit $?" >> $O/${T}_racecheck_fused_barrier.log
tail -3 $O/${T}_pytest.log; head -c 1500 $O/${T}_bench.json; echo; cat $O/${T}_bench_reference.json | head -c 600; echo
for f in $O/${T}_memcheck.log $O/${T}_racecheck_rest.log $O/${T}_racecheck_rest_lockstep.log $O/${T}_racecheck_fused.log $O/${T}_racecheck_fused_barrier.log; do echo "== $f"; tail -n 4 $f; done
ls -la $O | tail -25
