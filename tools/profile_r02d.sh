# Round-2 closing captures: the two attention kernels after their bulk-tensor-store epilogues, and the launch list.
set -x
export PCG_CUDA_GRAPHS=0
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --no-other-configs --no-full-last-block"
O=gpurun_out
$B > $O/r02d_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 --csv --log-file $O/r02d_launches.csv $B > $O/r02d_ncu0.log 2>&1
python tools/ncu_summ.py launches $O/r02d_launches.csv $O/r02d_launches_summary.csv "r02d: PCG_CUDA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 $B (ViT-L/14 x 128 cutouts, steady state: about two steps)"
rm -f $O/r02d_launches.csv
full() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $O/r02d_$1 $B > $O/r02d_ncu_$1.log 2>&1
  python tools/ncu_summ.py full $O/r02d_$1.ncu-rep $O/r02d_$1_ncu_full.csv "r02d: PCG_CUDA_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 $B"
}
full attn_fwd attn_fwd_split 30 2
python tools/ncu_stalls.py $O/r02d_attn_fwd.ncu-rep attn_fwd_split 25 > $O/r02d_attn_fwd_stalls.txt 2>&1
rm -f $O/r02d_attn_fwd.ncu-rep
full attn_bwd attn_bwd_persist 30 2
python tools/ncu_stalls.py $O/r02d_attn_bwd.ncu-rep attn_bwd_persist 25 > $O/r02d_attn_bwd_stalls.txt 2>&1
rm -f $O/r02d_attn_bwd.ncu-rep
ls -la $O/ | grep r02d
