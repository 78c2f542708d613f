import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        print("markers/s %.3e  step %.3f ms  phases %s  spread frac %.3f  interp frac %.3f  e2e %.3e" % (
            d["value"], d["ms_per_step"], {k: round(v, 3) for k, v in d["phases_ms"].items()}, d["roofline"]["frac"],
            d["roofline"]["interp"]["frac"], d["e2e"]["value"]))
    else:
        print(l[:300].rstrip())
