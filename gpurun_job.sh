timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not two_partitions" 2>&1 | tail -2
NSEM_KERNELS=v3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not two_partitions" 2>&1 | tail -2
timeout 900 python bench.py --cells 100 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
