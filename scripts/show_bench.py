"""print the key fields of bench.py JSON lines: python scripts/show_bench.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f)
    print("  value %.3f M f/s  ms/step %.4f  e2e %.3f M  e2e_model %s  launches %s  split_from %s" % (
        d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6,
        ("%.3f M" % (d["e2e_model"]["value"] / 1e6)) if "e2e_model" in d else None, d["gpu_launches"], d["config"].get("split_from")))
    r = d["roofline"]
    print("  parity %.3e  stage %s  executed %.1f TF (%.3f of %s)" % (d["parity"].get("max_abs_delta", -1), {k: round(v, 4) for k, v in r["stage_ms"].items()},
          r["executed_tflops"], r["executed_frac"], r["peak"]))
    print("  clocks", d.get("clocks"))
    for k, v in (d.get("secondary") or {}).items():
        print("  secondary %s: %.3f M f/s  ms/step %.4f e2e %.3f M parity %.3e stage %s" % (k, v["value"] / 1e6, v["ms_per_step"], v["e2e"] / 1e6,
              v["parity"]["max_abs_delta"], {a: round(b, 4) for a, b in v["stage_ms"].items()}))
    for v in d.get("variants") or []:
        print("  variant split_from %d: %.3f M f/s  ms/step %.4f  parity %.3e" % (v["split_from"], v["value"] / 1e6, v["ms_per_step"], v["parity_max_abs_delta"]))
    if d.get("cpu_baseline"):
        print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
