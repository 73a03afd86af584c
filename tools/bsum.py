"""stdin: bench.py output; prints a one-line summary (label = argv[1])."""
import json, sys
lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not lines:
    print(sys.argv[1] if len(sys.argv) > 1 else "", "NO JSON LINE")
    sys.exit(0)
d = json.loads(lines[-1])
k = d.get("kernel_ms_per_step", {})
print(sys.argv[1] if len(sys.argv) > 1 else "", "fps=%.0f ms/step=%.3f e2e=%.0f serial=%.3f |" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernel_timing_pass"]["ms_per_step"]),
      " ".join("%s=%.3f" % (n.replace("k_", ""), v) for n, v in k.items()))
