mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/kernel_timing.py --n 2 4 8 > gpurun_out/r2_sweep_small.jsonl 2>gpurun_out/r2_sweep_small.err; cut -c1-420 gpurun_out/r2_sweep_small.jsonl; tail -3 gpurun_out/r2_sweep_small.err
