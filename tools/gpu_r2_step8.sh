mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:vote_multi|numeric_direct_fast" -s 4 -c 2 -f -o gpurun_out/r2_small4 python tools/kernel_timing.py --n 4 --iters 3 > gpurun_out/ncu_small4.log 2>&1; tail -2 gpurun_out/ncu_small4.log
