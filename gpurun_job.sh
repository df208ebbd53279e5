timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r1_n1.json; cat gpurun_out/bench_r1_n1.json | cut -c1-3000
