N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_default_${N}gpu.json 2> gpurun_out/final_default_${N}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/final_default_${N}gpu.json').read().strip().splitlines()[-1]); print('default N=$N ms/step', d['ms_per_step'], 'value', d['value'], 'slab_bitwise', d['slab_bitwise'], 'clocks', d['clocks'], 'e2e', d['e2e']['value'], 'strong', d['strong'].get('ms_per_step'), d['strong'].get('value'), d['strong'].get('error'))"
