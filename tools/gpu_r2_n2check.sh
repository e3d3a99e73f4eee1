mkdir -p gpurun_out
NG=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NG --steps 100 --warmup 5 > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r2_bench_n$NG.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n$NG.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['compute_only'], d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')})
"; tail -3 gpurun_out/r2_bench_n$NG.err | cut -c1-300
