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
  dbg)
    timeout 200 python profiles/dbg_collinear.py > gpurun_out/r2/dbg_collinear.txt 2>&1; tail -40 gpurun_out/r2/dbg_collinear.txt | cut -c1-400 ;;
  quick)     # the tests that failed last time
    timeout 300 python -m pytest tests/test_trainer_trace_gpu.py tests/test_quirks_gpu.py -m gpu -q --timeout 200 > gpurun_out/r2/tests_quick.txt 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r2/tests_quick.txt; tail -30 gpurun_out/r2/tests_quick.txt | cut -c1-300
    timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "golden_exact" > gpurun_out/r2/tests_exact.txt 2>&1; echo "exact rc=$?" | tee -a gpurun_out/r2/tests_exact.txt; tail -30 gpurun_out/r2/tests_exact.txt | cut -c1-400 ;;
  bench2)    # what the driver runs at N = 2
    s=$(date +%s); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err; echo "bench2 rc=$? in $(( $(date +%s) - s )) s"
    tail -5 gpurun_out/r2/bench_n2.err | cut -c1-300; head -c 1500 gpurun_out/r2/bench_n2.json ;;
  mr2)
    timeout 400 python -m pytest tests/test_multirank_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2/tests_mr2.txt 2>&1; echo "mr2 rc=$?" | tee -a gpurun_out/r2/tests_mr2.txt; tail -20 gpurun_out/r2/tests_mr2.txt | cut -c1-300 ;;
  ptime)
    timeout 300 python profiles/policy_timing.py 262144 > gpurun_out/r2/policy_timing.txt 2>&1; cat gpurun_out/r2/policy_timing.txt | cut -c1-200 ;;
  ptime_ab)  # serial (UAISA_POLICY_PIPE=0) against pipelined forward
    for pp in ${PIPES:-0 7}; do echo "== UAISA_POLICY_PIPE=$pp"; FUSED_ONLY=1 UAISA_POLICY_PIPE=$pp timeout 200 python profiles/policy_timing.py 262144 2>&1 | grep -E "^----|fused" | cut -c1-200; done | tee gpurun_out/r2/policy_timing_ab.txt ;;
  ncu_policy)   # launch list + one full capture of the fused policy forward (sparse and mixed batches)
    for kind in sparse mixed; do
      timeout 200 python profiles/prof_policy.py $kind > gpurun_out/r2/prof_policy_$kind.log 2>&1 && \
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2/launches_policy_$kind.csv python profiles/prof_policy.py $kind > /dev/null 2>&1
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:uaisa_policy_forward -s 2 -c 1 -f -o gpurun_out/r2/prof_policy_$kind python profiles/prof_policy.py $kind > gpurun_out/r2/ncu_policy_$kind.log 2>&1; tail -2 gpurun_out/r2/ncu_policy_$kind.log
    done ;;
  ncu_env)      # one full capture of a warm 8-step rollout launch of the dense env workloads
    for w in c3 c4; do
      timeout 200 python profiles/prof_rollout.py $w 8 > gpurun_out/r2/prof_env_$w.log 2>&1 && \
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:uaisa_env -s 3 -c 1 -f -o gpurun_out/r2/prof_env_$w python profiles/prof_rollout.py $w 8 > gpurun_out/r2/ncu_env_$w.log 2>&1; tail -2 gpurun_out/r2/ncu_env_$w.log
    done ;;
  ncu_env_all)  # one full capture of a warm 8-step rollout launch per env workload (sat = few-neighbour warp-local kernel, c2 = owner / helper split, c3 / c4 dense)
    for w in sat c2 c3 c4; do
      timeout 200 python profiles/prof_rollout.py $w 8 > gpurun_out/r2/prof_env_$w.log 2>&1 && \
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:uaisa_env -s 3 -c 1 -f -o gpurun_out/r2/prof_env_$w python profiles/prof_rollout.py $w 8 > gpurun_out/r2/ncu_env_$w.log 2>&1; tail -1 gpurun_out/r2/ncu_env_$w.log
    done ;;
  launches)     # launch list of the default bench command (per-launch durations are cold-cache and serialised under ncu: shares, not absolutes)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:uaisa_ -c 400 --csv --log-file gpurun_out/r2/launches_bench_default.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2/ncu_bench.log 2>&1; tail -2 gpurun_out/r2/ncu_bench.log | cut -c1-200; wc -l gpurun_out/r2/launches_bench_default.csv ;;
  ncu_mlp)
    timeout 200 python profiles/prof_mlp.py > /dev/null 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:uaisa_mlp_gemm -s 8 -c 2 -f -o gpurun_out/r2/prof_mlp_gemm python profiles/prof_mlp.py > gpurun_out/r2/ncu_mlp.log 2>&1; tail -1 gpurun_out/r2/ncu_mlp.log ;;
  ppoprof)   # torch.profiler kernel tables of one update iteration and one rollout step
    timeout 300 python profiles/ppo_update_profile.py 65536 > gpurun_out/r2/ppo_update_profile_tc.txt 2>&1; head -16 gpurun_out/r2/ppo_update_profile_tc.txt | cut -c1-150
    timeout 300 python profiles/ppo_rollout_profile.py > gpurun_out/r2/ppo_rollout_profile.txt 2>&1; head -12 gpurun_out/r2/ppo_rollout_profile.txt | cut -c1-150 ;;
  ppo)
    timeout 600 python bench.py --workload ppo --ppo-epochs 3 > gpurun_out/r2/bench_ppo.json 2> gpurun_out/r2/bench_ppo.err; tail -3 gpurun_out/r2/bench_ppo.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/r2/bench_ppo.json') if l.startswith('{')][-1]);print(d['value']);print(d['ppo']['per_epoch'])" ;;
  fma)
    ./profiles/fma_peak > gpurun_out/r2/fma_peak.json 2> gpurun_out/r2/fma_peak.err; cat gpurun_out/r2/fma_peak.json ;;
  bench)
    s=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_default.json 2> gpurun_out/r2/bench_default.err; echo "bench rc=$? in $(( $(date +%s) - s )) s"
    tail -5 gpurun_out/r2/bench_default.err; head -c 2500 gpurun_out/r2/bench_default.json ;;
  ref)
    timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_reference.json 2> gpurun_out/r2/bench_reference.err; head -c 600 gpurun_out/r2/bench_reference.json ;;
esac
done
