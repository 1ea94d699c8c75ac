import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("img/s", round(d["value"]), "ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], " ".join(f"{k}={v['ms']}" for k,v in d["breakdown"].items()))
