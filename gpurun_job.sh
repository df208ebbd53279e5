for p in 0 1 2 3; do NSEM_PROBE=$p timeout 600 python probe_tmp.py 2>&1 | tail -1; done
