timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "steps_match_oracle or other_orders or vortex" 2>&1 | tail -3
timeout 300 python probe_tmp.py 2>&1 | tail -7
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v4_n100.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_v4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 6 -c 2 -o gpurun_out/prof_r1_v4_n100 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_v4.log 2>&1
tail -1 gpurun_out/ncu_full_v4.log | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('final', d['ms_per_step'], '%.3e'%d['value'], 'A %.2f ms %.3f'%(r['sweepA']['ms'], r['sweepA']['frac']), 'B %.2f ms %.3f'%(r['sweepB']['ms'], r['sweepB']['frac']), 'step %.3f'%r['step']['frac'], d['clocks'])"
