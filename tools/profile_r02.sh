# Round-2 profile captures (run under gpurun on one B200; see profiles/README.md).  Eager launches so that ncu sees
# every kernel; per-launch times are cold-cache and serialised: compare shares, not absolutes.
set -x
export PCG_CUDA_GRAPHS=0
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity"
O=gpurun_out
$B > $O/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 --csv --log-file $O/r02_launches.csv $B > $O/r02_ncu0.log 2>&1
python tools/ncu_summ.py launches $O/r02_launches.csv $O/r02_launches_summary.csv "r02: PCG_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 $B (ViT-L/14 x 128 cutouts, steady state: about two steps)"
rm -f $O/r02_launches.csv
full() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $O/r02_$1 $B > $O/r02_ncu_$1.log 2>&1
  python tools/ncu_summ.py full $O/r02_$1.ncu-rep $O/r02_$1_ncu_full.csv "r02: PCG_CUDA_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 $B"
}
full attn_fwd attn_fwd_persist 30 2
python tools/ncu_stalls.py $O/r02_attn_fwd.ncu-rep attn_fwd_persist 25 > $O/r02_attn_fwd_stalls.txt 2>&1
rm -f $O/r02_attn_fwd.ncu-rep
full attn_bwd attn_bwd_persist 30 2
python tools/ncu_stalls.py $O/r02_attn_bwd.ncu-rep attn_bwd_persist 25 > $O/r02_attn_bwd_stalls.txt 2>&1
rm -f $O/r02_attn_bwd.ncu-rep
full ln "layernorm_bwd|layernorm_fwd" 120 4
rm -f $O/r02_ln.ncu-rep
full sampler "sampler_fwd_vec|sampler_bwd_vec|head_|embed_" 14 9
rm -f $O/r02_sampler.ncu-rep
full gemm gemm_tcgen05 672 16
rm -f $O/r02_gemm.ncu-rep
ls -la $O/
