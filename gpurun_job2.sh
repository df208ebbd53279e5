timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 6 -c 2 -o gpurun_out/prof_v4c -f python bench.py --cells 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_v4c.log 2>&1
tail -2 gpurun_out/ncu_v4c.log | cut -c1-200
