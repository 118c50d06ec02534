#!/bin/bash
# A/B of libpcg.so variants on ONE box (boxes of the pool differ by +-3 %): tools/ab.sh ROUNDS name=path [name=path ...]
# ("intree" = perceptor_b200/libpcg.so; append ,VAR=val to set environment variables for that variant).  Each run: bench.py headline workload, 10 steps, no side legs; prints
# cutouts/s, ms/step and the per-family ms of the profiled pass.
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for v in "$@"; do
    name=${v%%=*}; rest=${v#*=}; path=${rest%%,*}; envs=""
    [ "$rest" != "$path" ] && envs=$(echo "${rest#*,}" | tr "," " ")   # name=path,VAR=val,VAR2=val2
    if [ "$path" = "intree" ]; then unset PCG_LIBRARY; else export PCG_LIBRARY=$path; fi
    env $envs timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-parity --no-other-configs --no-full-last-block ${AB_ARGS} > /tmp/ab.json 2>/tmp/ab.err \
      && python - "$name" <<'PY'
import json, sys
d = json.load(open("/tmp/ab.json"))
f = d["config"]["kernel_families"]
print(f"{sys.argv[1]:10s} {d['value']:8.1f} cut/s {d['ms_per_step']:7.3f} ms  " + " ".join(f"{k}={v['ms_per_step']:.2f}" for k, v in f.items()) + f"  clk={d['clocks']['sm_mhz']} W={d['clocks']['power_w_max']}", flush=True)
PY
    [ $? -ne 0 ] && { echo "$name FAILED"; tail -3 /tmp/ab.err; }
  done
done
