TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "pipelined or pipeline" > gpurun_out/t6_pipe.log 2>&1; echo "pipe test rc=$?"; tail -5 gpurun_out/t6_pipe.log
python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/t6_slabtest.log 2>&1; echo "slab test rc=$?"; tail -2 gpurun_out/t6_slabtest.log
for P in 1 0; do
  python bench.py --steps 100 --warmup 10 --no-extras --no-cpu-baseline --e2e-instances 1 --opt pipeline=$P > gpurun_out/t6_b1_p$P.json 2> gpurun_out/t6_b1_p$P.err
  $TR bench.py --gpus 2 --steps 100 --warmup 10 --no-extras --no-cpu-baseline --opt pipeline=$P > gpurun_out/t6_b2_p$P.json 2> gpurun_out/t6_b2_p$P.err
  python - <<P
import json
for n in ('b1','b2'):
    try:
        d=json.loads(open('gpurun_out/t6_%s_p$P.json'%n).read().strip().splitlines()[-1]); print(n,'pipeline $P ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'sweeps', d['roofline']['sweeps_psi'])
    except Exception as e: print(n,'pipeline $P failed', e)
P
done
python -m pytest tests -m gpu -x -q > gpurun_out/t6_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t6_tests.log
