"""Print a one-line summary of the last JSON line on stdin (bench.py output; NCCL may print its version before it)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
r = d["roofline"]
print(tag, d["config"].get("halo", ""), "N=%d ms/step %.3f A %.3f B %.3f bc %.3f %.3f value %.4g setup %.1fs" % (
    d["n_gpus"], d["ms_per_step"], r["sweepA"]["ms"], r["sweepB"]["ms"], r["bc_ms"][0], r["bc_ms"][1], d["value"], d["setup_s"]),
    "mp_parity", (d.get("mp_parity") or {}).get("bitwise"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
