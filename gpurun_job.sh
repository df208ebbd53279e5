set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --n 64 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
python bench.py --n 100 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
nproc; free -g | head -2
