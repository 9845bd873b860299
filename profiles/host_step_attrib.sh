# uaisa_step_host_sparse latency at configs[1]: single-kernel zero-copy path vs the copy-engine graph
B="python bench.py --no-cpu --no-sat --steps 200 --warmup 5"
echo zero-copy; UAISA_HOST_PROF=1 $B 2>&1 | grep "host step" | tail -1
echo copy-engine; UAISA_HOST_COPY_ENGINE=1 UAISA_HOST_PROF=1 $B 2>&1 | grep "host step" | tail -1
