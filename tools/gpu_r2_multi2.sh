set -x
NG=${1:-2}
mkdir -p gpurun_out
if [ "$NG" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_multi.txt; cat gpurun_out/r2_pytest_multi.txt
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29612 tools/check_reassembly.py --records 1000000 --sweep 1 --short 1 > gpurun_out/r2_push_sweep_n$NG.txt 2> gpurun_out/r2_push_sweep_n$NG.err; cat gpurun_out/r2_push_sweep_n$NG.txt; tail -5 gpurun_out/r2_push_sweep_n$NG.err
