import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(sys.argv[1] if len(sys.argv) > 1 else "", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"]["stage_ms_last_step"],
          {k.split(" ")[0]: round(v, 1) for k, v in d["roofline"]["families_ms"].items()}, "launches", d["gpu_launches"])
