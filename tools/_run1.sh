for P in 1 0; do
python bench.py --steps 100 --warmup 10 --no-extras --no-cpu-baseline --e2e-instances 1 --opt pdl=$P > gpurun_out/t7_pdl$P.json 2> gpurun_out/t7_pdl$P.err
python -c "
import json; d=json.loads(open('gpurun_out/t7_pdl$P.json').read().strip().splitlines()[-1]); print('pdl $P ms/step', d['ms_per_step'], 'launch us', d['roofline']['avg_launch_us'], d['roofline']['sweeps_psi'])"
done
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-instances 1 --workload cfg3 > gpurun_out/t7_cfg3.json 2> gpurun_out/t7_cfg3.err
python -c "
import json; d=json.loads(open('gpurun_out/t7_cfg3.json').read().strip().splitlines()[-1]); print('cfg3 ms/step', d['ms_per_step'], d['roofline']['sweeps_psi'], d['roofline']['sweeps_A'])"
python -m pytest tests -m gpu -x -q > gpurun_out/t7_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t7_tests.log
