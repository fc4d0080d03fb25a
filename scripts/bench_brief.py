"""Runs bench.py and prints a compact per-kernel summary (development aid)."""
import json, subprocess, sys, os
env = dict(os.environ)
out = subprocess.run([sys.executable, "bench.py", "--steps", "30"] + sys.argv[1:], capture_output=True, text=True, env=env).stdout.strip().splitlines()[-1]
d = json.loads(out)
print(f"{d['ms_per_step']:.4f} ms/step  {d['value']/1e6:.1f} Mpts/s  e2e {d['e2e']['value']/1e6:.1f}  frac {d['roofline_step']['frac']}")
print("  " + " | ".join(f"{k.replace('linear_','').replace('crf_','')}={v['ms_per_step']*1e3:.0f}" for k, v in d["kernels"].items()))
