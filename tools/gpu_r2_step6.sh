mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "weighted or config4" -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/config4_timing.py > gpurun_out/r2_config4.json 2> gpurun_out/r2_config4.err; cat gpurun_out/r2_config4.json; tail -3 gpurun_out/r2_config4.err
timeout 600 python tools/latency.py > gpurun_out/r2_latency.json 2> gpurun_out/r2_latency.err; cat gpurun_out/r2_latency.json; tail -5 gpurun_out/r2_latency.err
