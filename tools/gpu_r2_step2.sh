set -x
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; cut -c1-4000 gpurun_out/r2_bench_n1_a.json; tail -5 gpurun_out/r2_bench_n1_a.err
timeout 600 python bench.py --impl reference --steps 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err; cut -c1-1500 gpurun_out/r2_bench_ref_a.json; tail -3 gpurun_out/r2_bench_ref_a.err
