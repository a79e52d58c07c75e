import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d["degree"],d["dtype"],"stiff %.3f ms %.1f GDoF/s %.3f | mass %.3f ms %.3f"%(d["stiffness_ms"],d["stiffness_gdofs"],d["stiffness_frac"],d["mass_ms"],d["mass_frac"]))
