python -m pytest tests -m gpu -x -q > gpurun_out/t10_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t10_tests.log
python bench.py --steps 100 --warmup 10 --no-extras --no-cpu-baseline --e2e-instances 1 > gpurun_out/t10_b100.json 2> gpurun_out/t10_b100.err
python -c "
import json; d=json.loads(open('gpurun_out/t10_b100.json').read().strip().splitlines()[-1]); print('cfg2 100 steps ms/step', d['ms_per_step'], d['roofline']['avg_launch_us'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_cfg2.csv python bench.py --steps 10 --warmup 5 --no-extras --no-cpu-baseline --e2e-instances 1 > gpurun_out/t10_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_psi_tile -s 30 -c 2 -o gpurun_out/r02b_psi_tile_f32 -f python bench.py --steps 10 --warmup 5 --no-extras --no-cpu-baseline --e2e-instances 1 > gpurun_out/t10_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_psi_tile|k_a_tile" -s 10 -c 4 -o gpurun_out/r02b_tiles_f64 -f python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --e2e-instances 1 --workload cfg3 > gpurun_out/t10_ncu2.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/t10_default.json 2> gpurun_out/t10_default.err
python -c "
import json; d=json.loads(open('gpurun_out/t10_default.json').read().strip().splitlines()[-1]); print('default ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
