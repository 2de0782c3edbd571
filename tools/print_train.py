import json, sys
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print("ms/step", round(d["ms_per_step"], 3), "steps", d.get("step_ms"), "graph", d.get("cuda_graph"), "phases", d["phase_ms"],
              "launches", d["gpu_launches"], "loss", round(d["final_loss"], 5), "M samples/s", round(d["value"] / 1e6, 2))
