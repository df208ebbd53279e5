timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_gpu_check.py METIS 2>&1 | tail -2
for ov in 1 0; do for c in 64 100; do
NSEM_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --cells $c --steps 5 --warmup 3 --no-e2e --decomp METIS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('overlap $ov cells $c', d['n_gpus'], d['ms_per_step'], '%.3e'%d['value'], d['gpu_launches'], d['setup_s'])"
done; done
