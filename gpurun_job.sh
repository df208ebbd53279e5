timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1_n1.json; cat gpurun_out/bench_r1_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1
