import sys,json
for ln in sys.stdin.read().strip().splitlines():
    if ln.startswith('{'):
        d=json.loads(ln); r=d.get("roofline",{})
        print("value=%.0f Mpx/s  kernel=%.4f ms  frac=%.3f  achieved=%.0f GB/s  e2e=%.0f  launches=%s clocks=%s"%(d["value"], r.get("avg_launch_ms",0), r.get("frac",0), r.get("achieved",0), d["e2e"]["value"], d.get("gpu_launches"), d.get("clocks")))
    else: print(ln[:300])
