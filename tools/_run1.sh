python -m pytest tests -m gpu -x -q > gpurun_out/final_tests_1gpu.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/final_default_1gpu.json 2> gpurun_out/final_default_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/final_default_1gpu.json').read().strip().splitlines()[-1]); print('default ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'traffic', d['roofline']['traffic']); print({k:(v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), v.get('error')) for k,v in d['also'].items()})"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_reference_1gpu.json 2> gpurun_out/final_reference_1gpu.err; tail -c 400 gpurun_out/final_reference_1gpu.json
