TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/t4_slabtest.log 2>&1; echo "slab test rc=$?"; tail -2 gpurun_out/t4_slabtest.log
for B in 0 -1 8 32; do
  $TR bench.py --gpus 2 --steps 100 --warmup 10 --no-extras --no-cpu-baseline --opt slab_bnd=$B > gpurun_out/t4_b2_bnd$B.json 2> gpurun_out/t4_b2_bnd$B.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/t4_b2_bnd$B.json').read().strip().splitlines()[-1]); print('bnd $B ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
except Exception as e: print('bnd $B failed', e)
P
done
python bench.py --steps 100 --warmup 10 --no-extras --no-cpu-baseline --e2e-instances 1 > gpurun_out/t4_b1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/t4_b1.json').read().strip().splitlines()[-1]); print('1gpu ms/step', d['ms_per_step'])"
