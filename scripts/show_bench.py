import json, sys
p = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.json"
d = json.loads(open(p).read().strip().split("\n")[-1])
print("value %.0f frames/s  ms/step %.4f | e2e %.0f (%.4f ms) | stage_ms %s | frac %.3f exec %.1f TF/s | launches %d | clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"],
    d["roofline"]["frac"], d["roofline"].get("executed_tflops", 0), d["gpu_launches"], d["clocks"]))
