set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())"
timeout 900 python -m pytest tests/test_gpu_jsonpacked.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_jsonpacked.txt; cat gpurun_out/r2_pytest_jsonpacked.txt
timeout 600 python tools/jsonpacked_throughput.py --records 131072 --chunk-mb 16,64,128 --streams 1,3,4 > gpurun_out/r2_jsonpacked_sweep.jsonl 2> gpurun_out/r2_jsonpacked_sweep.err; tail -12 gpurun_out/r2_jsonpacked_sweep.jsonl | cut -c1-700; tail -3 gpurun_out/r2_jsonpacked_sweep.err
timeout 600 python tools/jsonpacked_throughput.py --records 1000000 --reps 3 > gpurun_out/r2_jsonpacked_1m.jsonl 2> gpurun_out/r2_jsonpacked_1m.err; cat gpurun_out/r2_jsonpacked_1m.jsonl | cut -c1-900; tail -3 gpurun_out/r2_jsonpacked_1m.err
timeout 900 python bench.py --steps 50 --warmup 5 --e2e-records 262144 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; cut -c1-3000 gpurun_out/r2_bench_n1_a.json; tail -5 gpurun_out/r2_bench_n1_a.err
