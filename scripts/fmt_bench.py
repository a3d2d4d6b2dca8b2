import sys, json
for path in sys.argv[1:]:
    for l in open(path):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline", {})
            print(path, "n_gpus", d.get("n_gpus"), "value", round(d["value"] / 1e6, 1), "M docs/s", "ms/step", round(d["ms_per_step"], 3),
                  "kernel_ms", round(r.get("kernel_ms", 0), 3), "GB/s", round(r.get("achieved", 0)), "frac", round(r.get("frac", 0), 3),
                  "e2e", round(d.get("e2e", {}).get("value", 0) / 1e6, 1), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
