"""Run bench.py over psi-kernel variants and print one summary line each (GPU box helper)."""
import json
import subprocess
import sys

variants = sys.argv[1:] or ["0:1", "1:2", "1:3", "1:4", "1:5", "1:6"]
extra = []
if "--" in variants:
    i = variants.index("--")
    variants, extra = variants[:i], variants[i + 1:]
for v in variants:
    pk, k = v.split(":")
    out = subprocess.run([sys.executable, "bench.py", "--psi-kernel", pk, "--psi-k", k, "--no-cpu-baseline"] + extra,
                         capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else ""
    try:
        d = json.loads(line)
        r = d["roofline"]
        print("kernel %s K %s: %.3e cell-steps/s  %.3f ms/step  frac %.3f  launches %d avg %.1f us  sweeps %d/%d replays %s  e2e %.3e  clocks %s"
              % (pk, k, d["value"], d["ms_per_step"], r["frac"], d["gpu_launches"], r["avg_launch_us"], r["sweeps_psi"],
                 r["sweeps_A"], d["replays"], d["e2e"]["value"], d["clocks"]["sm_mhz"]), flush=True)
    except Exception as e:
        print("kernel %s K %s FAILED: %s\n%s\n%s" % (pk, k, e, out.stdout[-500:], out.stderr[-1500:]), flush=True)
