mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/kernel_timing.py --n 2 4 8 16 32 64 > gpurun_out/r2_sweep_kernels_n1.jsonl 2>gpurun_out/r2_sweep_kernels_n1.err; cat gpurun_out/r2_sweep_kernels_n1.jsonl; tail -3 gpurun_out/r2_sweep_kernels_n1.err
KC_VOTE_MULTI=0 timeout 600 python tools/kernel_timing.py --n 2 4 8 2>/dev/null | cut -c1-200
