import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g e2e %.4g kernel_ms %.5f frac %.4f"%(d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms"],d["roofline"]["frac"]))
if "step_form" in d: print("step_form",d["step_form"]["ms_per_launch"],d["step_form"]["roofline"]["frac"])
if "mpc" in d: print("mpc",d["mpc"]["value"],d["mpc"]["episode"])
print(d["qoe_stats"]["reward"], d["flagged_sessions"])
