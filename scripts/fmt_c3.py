import sys, json
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        print(tag, d["config"][:24], "ms", round(d["kernel_ms"], 1), "TF", round(d["useful_TFLOPs"]), "MHz", d.get("sm_mhz"), "W", d.get("power_w"),
              "clkfrac", round(d.get("frac_of_clock_peak") or 0, 3), "sust", round(d["frac_tensor_sustained"], 3))
    elif "Error" in l or "error" in l:
        print(tag, l.strip()[:300])
