# Round-2 last capture: one layer's GEMMs with the bulk-tensor-store epilogue (see tools/profile_r02c.sh for the windows).
set -x
export PCG_CUDA_GRAPHS=0
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --no-other-configs --no-full-last-block"
O=gpurun_out
full() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $O/r02e_$1 $B > $O/r02e_ncu_$1.log 2>&1
  python tools/ncu_summ.py full $O/r02e_$1.ncu-rep $O/r02e_$1_ncu_full.csv "r02e: PCG_CUDA_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 $B"
  rm -f $O/r02e_$1.ncu-rep
}
full gemm_fwd gemm_tcgen05 626 4
full gemm_bwd gemm_tcgen05 707 4
