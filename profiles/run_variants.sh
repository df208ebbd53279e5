#!/bin/bash
# run_variants.sh <cells> <name>...: time each variant build back to back on one box (bench.py, device-resident, no CPU baseline / e2e)
cells=$1; shift
for v in "$@"; do
  NSEM_LIBDIR=variants/$v python bench.py --cells $cells --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/var_$v.err | python profiles/jsonline.py "$v"
done
