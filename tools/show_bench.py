"""Pretty-print the one-line JSON of bench.py (stdin)."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.0f frames/s | ms/step %.1f | e2e %s | launches %s" % (
    d["value"], d["ms_per_step"], ("%.0f" % d["e2e"]["value"]) if d.get("e2e") else None, d["gpu_launches"]))
r = d["roofline"]
print("dominant %s achieved %.1f TFLOP/s avg %.3f ms frac %.3f | whole-step %.1f TFLOP/s" % (
    r["kernel"], r["achieved"], r["avg_launch_ms"], r["frac"], r["whole_step_tflops"]))
print("shares", r["time_shares"])
print("clocks", d["clocks"])
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
if d.get("torch_gpu_reference"): print("torch-gpu-port", d["torch_gpu_reference"])
