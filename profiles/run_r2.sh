#!/bin/bash
# Round-2 GPU session driver (run under gpurun from the repo root): GPU tests, the default bench line, the reference arm, the FMA peak.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/run_r2.sh [tests|bench|all]'
what=${1:-all}
mkdir -p gpurun_out/r2
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2/tests.txt
  tail -15 gpurun_out/r2/tests.txt
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  ./profiles/fma_peak > gpurun_out/r2/fma_peak.json 2> gpurun_out/r2/fma_peak.err; cat gpurun_out/r2/fma_peak.json
  /usr/bin/time -v timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_default.json 2> gpurun_out/r2/bench_default.err; echo "bench rc=$?"
  tail -3 gpurun_out/r2/bench_default.err; head -c 3000 gpurun_out/r2/bench_default.json
  timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_reference.json 2> gpurun_out/r2/bench_reference.err
fi
