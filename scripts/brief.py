import json, sys
d = json.loads(sys.stdin.read())
print(sys.argv[1:] , d["value"], d["ms_per_step"], "upd/solve host ms", d.get("update_ms_host_api"), d.get("solve_ms_host_api"), "e2e", d["e2e"]["value"], "upd TF", d["roofline"]["achieved"],
      {k: v["ms"] for k, v in d["phases_one_step"].items() if v["ms"] > 0.05})
