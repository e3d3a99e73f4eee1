# Round-2 recipe: gpurun --gpus N -- bash tools/run_gpu_sweep_multi.sh N   (bench.py --sweep over n at N GPUs)
NG=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $NG --sweep 2,4,8,16,32,64 --steps 50 --warmup 5 > gpurun_out/r2_sweep_n$NG.jsonl 2> gpurun_out/r2_sweep_n$NG.err; python -c "
import json
for l in open('gpurun_out/r2_sweep_n$NG.jsonl'):
    d=json.loads(l); r=d['roofline']; print('n='+d['config']['workload'].split('n=')[1][:3], round(d['value']/1e9,3),'G rec/s', round(d['ms_per_step'],4),'ms', 'compute_only', round(d['compute_only']['value']/1e9,2), d['compute_only']['reassembly'], 'floor', round(d['compute_only']['nvlink_floor_ms'],3))
"; tail -3 gpurun_out/r2_sweep_n$NG.err | cut -c1-300
