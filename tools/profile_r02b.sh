set -x
export PCG_CUDA_GRAPHS=0
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity"
O=gpurun_out
$B > $O/r02b_plain.log 2>&1 || exit 1
full() {
  ncu --set full --clock-control none -k regex:"$2" -s $3 -c $4 -o $O/r02_$1 $B > $O/r02_ncu_$1.log 2>&1
  python tools/ncu_summ.py full $O/r02_$1.ncu-rep $O/r02_$1_ncu_full.csv "r02: PCG_CUDA_GRAPHS=0 ncu --set full --clock-control none -k regex:$2 -s $3 -c $4 $B"
  rm -f $O/r02_$1.ncu-rep
}
full ln_bwd layernorm_bwd 100 2
full sampler_only "sampler_fwd_vec|sampler_bwd_vec" 4 2
ls -la $O
