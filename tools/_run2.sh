N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for SW in "20 5" "100 10"; do set -- $SW
$TR bench.py --gpus $N --steps $1 --warmup $2 --no-extras --no-cpu-baseline > gpurun_out/t9_w${N}_$1.json 2> gpurun_out/t9_w${N}_$1.err
python -c "
import json; d=json.loads(open('gpurun_out/t9_w${N}_$1.json').read().strip().splitlines()[-1]); print('weak N=$N steps $1 ms/step', d['ms_per_step'], 'value', d['value'], d['clocks'])"
done
