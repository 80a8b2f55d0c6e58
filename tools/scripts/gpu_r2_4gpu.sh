# 4-GPU validation of the bench contract: torchrun, barrier + max over ranks, configs block, strong C4
set -x
O=gpurun_out/r2_4gpu
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 3 > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "rc=$?" >> $O/bench_4gpu.err
tail -3 $O/bench_4gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > $O/bench_ref_4gpu.json 2> $O/bench_ref_4gpu.err; echo "rc=$?" >> $O/bench_ref_4gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_4gpu/bench_4gpu.json'))
print('n_gpus',d['n_gpus'],'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'], 'u8 %.4g'%d['e2e']['pcm16_in_u8_out']['value'])
print({k:(round(v.get('value',0)),v.get('scaling')) for k,v in d['configs'].items()})
r=json.load(open('gpurun_out/r2_4gpu/bench_ref_4gpu.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
