TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== 1 GPU"; python tools/diag_chunks.py 2>&1 | grep chunk
echo "== 2 GPUs"; $TR tools/diag_chunks.py 2>&1 | grep chunk
echo "== 2 GPUs pipeline=0"; $TR tools/diag_chunks.py pipeline=0 2>&1 | grep chunk
