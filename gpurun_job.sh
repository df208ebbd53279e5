timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "steps_match_oracle or other_orders" 2>&1 | tail -3
B='timeout 600 python bench.py --cells 100 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e'
P='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(sys.argv[1], d["ms_per_step"], "%.3e"%d["value"], "A %.2f ms"%(r["sweepA"]["ms"]), "B %.2f ms"%(r["sweepB"]["ms"]), r["bc_ms"])'
$B 2>&1 | tail -1 | python -c "$P" v4c
