import json,sys
d=json.loads(open(sys.argv[1]).readline())
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks","n_gpus")})
print("frac", d["config"]["frac_of_measured_bf16_peak"], "launch", d["config"]["launch"])
if d.get("roofline"): print("roof", d["roofline"]["frac"], d["roofline"]["fwd_us"], d["roofline"]["bwd_us"])
print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"])
for k,v in d["extra"].items():
    print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ("note","layers")})
