#!/bin/bash
# run_variants.sh <cells> <name>...: time each variant build back to back on one box (bench.py, device-resident, no CPU baseline / e2e)
cells=$1; shift
for v in "$@"; do
  NSEM_LIBDIR=variants/$v python bench.py --cells $cells --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/var_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v', 'cells', $cells, 'ms/step %.3f A %.3f B %.3f bc %.3f %.3f' % (d['ms_per_step'], r['sweepA']['ms'], r['sweepB']['ms'], r['bc_ms'][0], r['bc_ms'][1]), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
