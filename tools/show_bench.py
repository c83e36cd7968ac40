"""Condense bench.py's JSON line (stdin) to the figures worth eyeballing."""
import json, sys
d = json.loads([l for l in sys.stdin if l.startswith("{")][-1])
e = d.get("e2e") or {}
print("N=%d ms/step %.4f value %.3e e2e %.3e e2e-iters %s clocks %s" % (
    d["n_gpus"], d["ms_per_step"], d["value"], e.get("value", 0), e.get("iteration_ms_rank0"), d["clocks"]))
