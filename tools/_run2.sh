TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SVL_PDL=2 timeout 300 python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/t11_slabtest.log 2>&1; echo "slab test (pdl=2) rc=$?"; tail -2 gpurun_out/t11_slabtest.log
for P in 2 1; do
  SVL_PDL=$P timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 10 --no-extras --no-cpu-baseline > gpurun_out/t11_b2_pdl$P.json 2> gpurun_out/t11_b2_pdl$P.err
  python -c "
import json; d=json.loads(open('gpurun_out/t11_b2_pdl$P.json').read().strip().splitlines()[-1]); print('2 GPUs pdl $P ms/step', d['ms_per_step'])"
done
for SH in 1 2; do
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-instances 1 --workload cfg3 --opt psi_shape=$SH > gpurun_out/t11_cfg3_sh$SH.json 2> gpurun_out/t11_cfg3_sh$SH.err
python -c "
import json; d=json.loads(open('gpurun_out/t11_cfg3_sh$SH.json').read().strip().splitlines()[-1]); print('cfg3 shape $SH ms/step', d['ms_per_step'])"
done
ncu --set full --clock-control none --import-source on -k regex:k_psi_tile -s 8 -c 1 -o gpurun_out/r02b_psi_tile_f64 -f python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --e2e-instances 1 --workload cfg3 > gpurun_out/t11_ncu.log 2>&1
