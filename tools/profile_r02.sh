set -x
export PCG_CUDA_GRAPHS=0
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity"
$B > gpurun_out/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 800 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_persist -s 30 -c 2 -o gpurun_out/r02_attn_fwd $B > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_persist -s 30 -c 2 -o gpurun_out/r02_attn_bwd $B > gpurun_out/r02_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"layernorm_bwd|layernorm_fwd" -s 120 -c 4 -o gpurun_out/r02_ln $B > gpurun_out/r02_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:"sampler_fwd_vec|sampler_bwd_vec|head_|embed_" -s 14 -c 9 -o gpurun_out/r02_sampler $B > gpurun_out/r02_ncu4.log 2>&1
ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 672 -c 16 -o gpurun_out/r02_gemm $B > gpurun_out/r02_ncu5.log 2>&1
ls -la gpurun_out/r02_*
