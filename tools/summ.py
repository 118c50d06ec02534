import json, sys
for f in sys.argv[1:]:
    lines=[l for l in open(f) if l.startswith("{")]
    if not lines: print(f, open(f).read()[-2000:]); continue
    d=json.loads(lines[-1])
    print(f, "value",round(d["value"],1),"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],1),"launches",d["gpu_launches"], "stepfrac", round(d["config"]["step_frac_of_peak"],3))
    r=d["roofline"]; print("  gemm TF", round(r["achieved"],1), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3))
    for k,v in d["config"]["kernel_families"].items(): print("   ",k,{a:round(b,3) for a,b in v.items()})
    print("  clocks", d["clocks"])
