# ncu --set full of the depth backbone's GEMM-class kernels (first training step of tools/depth_bench.py --profile)
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|gemm_tn2_kernel|gemm_tc2_kernel" -c 52 -f -o /tmp/prof_depth \
   python tools/depth_bench.py --profile > gpurun_out/depth_ncu_full.out 2>&1; echo "rc=$?"
python tools/ncu_summary.py report /tmp/prof_depth.ncu-rep gpurun_out/depth_ncu_full.txt
ls -la /tmp/prof_depth.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"bn_bwd_apply_kernel|bn_apply_kernel|bn_reduce_kernel|col2im_kernel|im2col_kernel|maxpool" -c 16 -f -o /tmp/prof_depth_bn \
   python tools/depth_bench.py --profile > gpurun_out/depth_ncu_full_bn.out 2>&1; echo "rc=$?"
python tools/ncu_summary.py report /tmp/prof_depth_bn.ncu-rep gpurun_out/depth_ncu_full_bn.txt
