# Final round-2 captures of the kernels that changed after profile_r02.sh (run under gpurun on one B200).  Eager
# launches so that ncu sees every kernel; per-launch times are cold-cache and serialised: compare shares, not absolutes.
set -x
export PCG_CUDA_GRAPHS=0
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --no-other-configs --no-full-last-block"
O=gpurun_out
$B > $O/r02c_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 --csv --log-file $O/r02c_launches.csv $B > $O/r02c_ncu0.log 2>&1
python tools/ncu_summ.py launches $O/r02c_launches.csv $O/r02c_launches_summary.csv "r02c: PCG_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 $B (ViT-L/14 x 128 cutouts, steady state: about two steps)"
rm -f $O/r02c_launches.csv
full() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $O/r02c_$1 $B > $O/r02c_ncu_$1.log 2>&1
  python tools/ncu_summ.py full $O/r02c_$1.ncu-rep $O/r02c_$1_ncu_full.csv "r02c: PCG_CUDA_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 $B"
}
full attn_fwd attn_fwd_split 30 2
python tools/ncu_stalls.py $O/r02c_attn_fwd.ncu-rep attn_fwd_split 25 > $O/r02c_attn_fwd_stalls.txt 2>&1
rm -f $O/r02c_attn_fwd.ncu-rep
full attn_cls attn_cls 2 2
rm -f $O/r02c_attn_cls.ncu-rep
full ln "layernorm_bwd|layernorm_fwd" 120 4
rm -f $O/r02c_ln.ncu-rep
# one layer's GEMMs in the timed (4th) step: 195 GEMM launches per step = 98 forward (patch, 23 x {qkv, out, fc, proj}, last
# block: kv, q, out, fc, proj on the class-token rows) + 97 backward (last block: 3 class-token GEMMs + dqkv, 23 x {dproj, dfc,
# dout, dqkv}, dpatch).  626 = 3 * 195 + 1 + 4 * 10: forward layers 10 and 11; 707 = 3 * 195 + 98 + 4 + 4 * 5: backward layers 17, 16
full gemm_fwd gemm_tcgen05 626 8
rm -f $O/r02c_gemm_fwd.ncu-rep
full gemm_bwd gemm_tcgen05 707 8
rm -f $O/r02c_gemm_bwd.ncu-rep
ls -la $O/ | grep r02c
