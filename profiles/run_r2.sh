#!/bin/bash
# Round-2 GPU session driver (run under gpurun from the repo root).  Every step has its own timeout so that a hung kernel costs seconds, not the call.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/run_r2.sh <step> [<step> ...]'
# steps: newtests | fwd | alltests | bench | ref | fma
mkdir -p gpurun_out/r2
for what in "$@"; do
case $what in
  fwd)       # the full fused forward kernel, alone, short leash
    timeout 240 python -m pytest tests/test_policy.py -m gpu -x -q --timeout 120 > gpurun_out/r2/tests_fwd.txt 2>&1; echo "fwd rc=$?" | tee -a gpurun_out/r2/tests_fwd.txt; tail -25 gpurun_out/r2/tests_fwd.txt ;;
  newtests)
    for f in test_multirank_gpu test_ppo_gpu test_quirks_gpu test_rvo_select_gpu test_trainer_trace_gpu test_post_train_gpu test_mdin_adapter_gpu; do
      timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 280 > gpurun_out/r2/tests_$f.txt 2>&1; echo "$f rc=$?" | tee -a gpurun_out/r2/tests_$f.txt; tail -30 gpurun_out/r2/tests_$f.txt | cut -c1-300
    done ;;
  alltests)
    timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/r2/tests.txt 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2/tests.txt; tail -15 gpurun_out/r2/tests.txt ;;
  fma)
    ./profiles/fma_peak > gpurun_out/r2/fma_peak.json 2> gpurun_out/r2/fma_peak.err; cat gpurun_out/r2/fma_peak.json ;;
  bench)
    s=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_default.json 2> gpurun_out/r2/bench_default.err; echo "bench rc=$? in $(( $(date +%s) - s )) s"
    tail -5 gpurun_out/r2/bench_default.err; head -c 2500 gpurun_out/r2/bench_default.json ;;
  ref)
    timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_reference.json 2> gpurun_out/r2/bench_reference.err; head -c 600 gpurun_out/r2/bench_reference.json ;;
esac
done
