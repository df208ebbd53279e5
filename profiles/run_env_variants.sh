#!/bin/bash
# run_env_variants.sh <cells> <variant> <ENVVAR> <value>...: the same build timed with different values of one environment variable
cells=$1; v=$2; var=$3; shift 3
for val in "$@"; do
  env $var=$val NSEM_LIBDIR=variants/$v python bench.py --cells $cells --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/var_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v $var=$val', 'cells', $cells, 'ms/step %.3f A %.3f B %.3f bc %.3f %.3f' % (d['ms_per_step'], r['sweepA']['ms'], r['sweepB']['ms'], r['bc_ms'][0], r['bc_ms'][1]), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
