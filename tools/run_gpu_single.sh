# Round-2 recipe (run under gpurun on a B200 box): every GPU test, smoke, bench, sweep, config 3/4 and latency tools -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu_all.txt; cat gpurun_out/r2_pytest_gpu_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-600 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --sweep 2,4,8,16,32,64 --steps 50 --warmup 5 > gpurun_out/r2_sweep_bench_n1.jsonl 2> gpurun_out/r2_sweep_bench_n1.err; python -c "
import json
for l in open('gpurun_out/r2_sweep_bench_n1.jsonl'):
    d=json.loads(l); r=d['roofline']; print('n='+d['config']['workload'].split('n=')[1][:3], round(d['value']/1e9,3),'G rec/s', round(d['ms_per_step'],4),'ms', {k:round(v['frac'],3) for k,v in r['all_kernels'].items()}, 'step', round(r['step_frac'],3))
"
timeout 300 python tools/config4_timing.py --min-tokens 8 --max-tokens 64 > gpurun_out/r2_config4.json; cat gpurun_out/r2_config4.json
timeout 600 python tools/config3_timing.py --records 20000 > gpurun_out/r2_config3.json 2> gpurun_out/r2_config3.err; cut -c1-700 gpurun_out/r2_config3.json; tail -3 gpurun_out/r2_config3.err
timeout 600 python tools/latency.py > gpurun_out/r2_latency.json 2> gpurun_out/r2_latency.err; cat gpurun_out/r2_latency.json
timeout 300 python tools/jsonpacked_throughput.py --workload invoice --records 131072 --n 8 --reps 3 > gpurun_out/r2_jsonpacked_invoice_n8.json; cut -c1-400 gpurun_out/r2_jsonpacked_invoice_n8.json
timeout 300 python tools/jsonpacked_throughput.py --workload invoice_nested --records 131072 --n 8 --reps 3 > gpurun_out/r2_jsonpacked_invoice_nested_n8.json; cut -c1-400 gpurun_out/r2_jsonpacked_invoice_nested_n8.json
timeout 300 python tools/config4_timing.py --iters 30 > gpurun_out/r2_config4_k3b_tma_32_96.json; cat gpurun_out/r2_config4_k3b_tma_32_96.json
KC_K3B_ROWS=0 timeout 300 python tools/config4_timing.py --iters 30; KC_K3B_TMA=0 timeout 300 python tools/config4_timing.py --iters 30   # K3b: in-loop weights; one row per thread from global memory
