set -x
NG=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29612 tools/check_reassembly.py --records 1000000 --sweep 1 --short 1 > gpurun_out/r2_push_sweep_n$NG.txt 2> gpurun_out/r2_push_sweep_n$NG.err; cat gpurun_out/r2_push_sweep_n$NG.txt; tail -5 gpurun_out/r2_push_sweep_n$NG.err
for ch in 0; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NG --steps 50 --warmup 5 --e2e-records 262144 --e2e-columnar 0 --push-chunks $ch > gpurun_out/r2_bench_n${NG}_hybrid$ch.json 2> gpurun_out/r2_bench_n${NG}_hybrid$ch.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r2_bench_n${NG}_hybrid$ch.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['compute_only'], d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')})
"; tail -3 gpurun_out/r2_bench_n${NG}_hybrid$ch.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NG --steps 50 --warmup 5 --no-e2e --route peers > gpurun_out/r2_bench_n${NG}_peers.json 2> gpurun_out/r2_bench_n${NG}_peers.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r2_bench_n${NG}_peers.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['compute_only'])"
