set -x
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -12
if [ "$NG" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_multi.txt; cat gpurun_out/r2_pytest_multi.txt
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29612 tools/check_reassembly.py --records 1000000 > gpurun_out/r2_check_reassembly_n$NG.txt 2> gpurun_out/r2_check_reassembly_n$NG.err; cat gpurun_out/r2_check_reassembly_n$NG.txt; tail -5 gpurun_out/r2_check_reassembly_n$NG.err
for ch in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NG --steps 50 --warmup 5 --e2e-records 262144 --e2e-columnar 0 --push-chunks $ch > gpurun_out/r2_bench_n${NG}_push$ch.json 2> gpurun_out/r2_bench_n${NG}_push$ch.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r2_bench_n${NG}_push$ch.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['compute_only'], d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')})
"; tail -3 gpurun_out/r2_bench_n${NG}_push$ch.err
done
