import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json, subprocess
out = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--no-f32", "--steps", "10"], capture_output=True, text=True).stdout
l = json.loads(out.strip().splitlines()[-1])
print("ms/step %.4f  K1 %.4f  K2 %.4f  stats %.4f  frac %.4f" % (l["ms_per_step"], l["kernels"]["zstep_k1_ms"], l["kernels"]["zsolve_k2_ms"], l["kernels"]["stats_ms"], l["roofline"]["frac"]))
