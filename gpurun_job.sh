set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
NSEM_KERNELS=v1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bubble3d" 2>&1 | tail -2
timeout 600 python bench.py --n 64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['sweepA'], d['roofline']['sweepB'], d['roofline']['step'])"
timeout 600 python bench.py --n 100 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['sweepA'], d['roofline']['sweepB'], d['roofline']['step'])"
