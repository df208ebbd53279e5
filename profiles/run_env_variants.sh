#!/bin/bash
# run_env_variants.sh <cells> <variant> <ENVVAR> <value>...: the same build timed with different values of one environment variable
cells=$1; v=$2; var=$3; shift 3
for val in "$@"; do
  env $var=$val NSEM_LIBDIR=variants/$v python bench.py --cells $cells --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/var_$v.err | python profiles/jsonline.py "$v $var=$val"
done
