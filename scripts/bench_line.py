"""stdin: one bench.py JSON line -> a one-line summary (used by scripts/gpu_matrix.sh)"""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read())
c = d["config"]
print(tag, "ms/step", round(d["ms_per_step"], 3), "fanout", round(c["fanout_ms"], 3), "render", round(c.get("render_ms", 0), 3),
      "direct", round(c["direct_ms"], 3), "plan", round(c["plan_ms"], 3), "frac", round(d["roofline"]["frac"], 3),
      "| timed:", {k: round(v, 3) for k, v in c.get("timed_region_ms", {}).items() if k != "how"})
