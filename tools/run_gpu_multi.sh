# Round-2 recipe: gpurun --gpus N -- bash tools/run_gpu_multi.sh N   (cross-rank checks at N = 2, bench at N, config-5 sweep at N = 8)
NG=${1:-2}
mkdir -p gpurun_out
if [ "$NG" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_pytest_multi.txt; cat gpurun_out/r2_pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29612 tools/check_reassembly.py --records 1000000 > gpurun_out/r2_check_reassembly_n2.txt 2> gpurun_out/r2_check_reassembly_n2.err; cat gpurun_out/r2_check_reassembly_n2.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NG > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r2_bench_n$NG.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n$NG.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['compute_only'], d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')}, d['clocks'])
"; tail -3 gpurun_out/r2_bench_n$NG.err
if [ "$NG" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --sweep 2,4,8,16,32,64 --steps 50 --warmup 5 > gpurun_out/r2_sweep_n8.jsonl 2> gpurun_out/r2_sweep_n8.err; python -c "
import json
for l in open('gpurun_out/r2_sweep_n8.jsonl'):
    d=json.loads(l); r=d['roofline']; print(d['config']['workload'].split('n=')[1][:3], round(d['value']/1e9,3),'G rec/s', round(d['ms_per_step'],4),'ms', 'compute_only', round(d['compute_only']['value']/1e9,2), d['compute_only']['reassembly'])
"; tail -3 gpurun_out/r2_sweep_n8.err
fi
