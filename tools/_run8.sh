N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 100 --warmup 10 --no-extras --no-cpu-baseline > gpurun_out/t8_w$N.json 2> gpurun_out/t8_w$N.err
python -c "
import json; d=json.loads(open('gpurun_out/t8_w$N.json').read().strip().splitlines()[-1]); print('weak N=$N ms/step', d['ms_per_step'], 'value', d['value'])"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/t8_default$N.json 2> gpurun_out/t8_default$N.err
python -c "
import json; d=json.loads(open('gpurun_out/t8_default$N.json').read().strip().splitlines()[-1]); print('default N=$N ms/step', d['ms_per_step'], 'value', d['value'], 'slab_bitwise', d['slab_bitwise'], 'strong', d['strong'].get('ms_per_step'), d['strong'].get('value'), d['strong'].get('error'))"
