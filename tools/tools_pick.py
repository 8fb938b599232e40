import sys, json
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "value", "%.4g" % d["value"], "frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
