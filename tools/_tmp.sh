mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "weighted or config4" 2>&1 | tail -6 > gpurun_out/r2i_pytest_k3b.txt; cat gpurun_out/r2i_pytest_k3b.txt
timeout 200 python tools/config4_timing.py --min-tokens 8 --max-tokens 64 > gpurun_out/r2i_config4_minb3.json 2> gpurun_out/r2i.err; cat gpurun_out/r2i_config4_minb3.json; tail -3 gpurun_out/r2i.err
KC_K3B_MINB=2 timeout 200 python tools/config4_timing.py --min-tokens 8 --max-tokens 64 > gpurun_out/r2i_config4_minb2.json 2>> gpurun_out/r2i.err; cat gpurun_out/r2i_config4_minb2.json
timeout 200 python tools/config4_timing.py --min-tokens 8 --max-tokens 64 --n 64 --records 131072 > gpurun_out/r2i_config4_n64.json 2>> gpurun_out/r2i.err; cat gpurun_out/r2i_config4_n64.json
KC_K3B_PRE=0 timeout 200 python tools/config4_timing.py --min-tokens 8 --max-tokens 64 --n 64 --records 131072 > gpurun_out/r2i_config4_n64_nopre.json 2>> gpurun_out/r2i.err; cat gpurun_out/r2i_config4_n64_nopre.json
