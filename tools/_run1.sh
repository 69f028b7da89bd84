for SH in 0 1; do
python bench.py --steps 100 --warmup 10 --no-extras --no-cpu-baseline --e2e-instances 1 --opt psi_shape=$SH > gpurun_out/t5_sh$SH.json 2> gpurun_out/t5_sh$SH.err
python -c "
import json; d=json.loads(open('gpurun_out/t5_sh$SH.json').read().strip().splitlines()[-1]); print('shape $SH ms/step', d['ms_per_step'], 'launch us', d['roofline']['avg_launch_us'])"
done
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-instances 1 --workload cfg3 > gpurun_out/t5_cfg3.json 2> gpurun_out/t5_cfg3.err
python -c "
import json; d=json.loads(open('gpurun_out/t5_cfg3.json').read().strip().splitlines()[-1]); print('cfg3 ms/step', d['ms_per_step'], d['roofline']['sweeps_psi'], d['roofline']['sweeps_A'])"
python bench.py --steps 10 --warmup 3 --workload cfg5 > gpurun_out/t5_cfg5.json 2> gpurun_out/t5_cfg5.err
python -c "
import json; d=json.loads(open('gpurun_out/t5_cfg5.json').read().strip().splitlines()[-1]); print('cfg5 ms/step', d['ms_per_step'], d.get('vortices'))"
python -m pytest tests -m gpu -x -q > gpurun_out/t5_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t5_tests.log
