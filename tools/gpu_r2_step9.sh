mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/kernel_timing.py --n 64 > gpurun_out/r2_sweep_64.jsonl 2>gpurun_out/r2_sweep_64.err; cut -c1-420 gpurun_out/r2_sweep_64.jsonl; tail -3 gpurun_out/r2_sweep_64.err
KC_NUM_WIDE=0 timeout 600 python tools/kernel_timing.py --n 64 2>/dev/null | cut -c1-300
timeout 600 python tools/kernel_timing.py --n 48 2>/dev/null | cut -c1-300
