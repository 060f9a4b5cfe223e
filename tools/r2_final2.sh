# round 2 record run (final build) on one B200: both bench arms, ncu launch list + full captures summarised on the box
set -x
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -9
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 300 gpurun_out/r2_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv \
  python tools/infer_probe.py --images 4 --once --precision f16c8 > gpurun_out/ncu_l.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_launches.csv gpurun_out/r2_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'pairs_|tokens_kernel|postprocess_kernel|obj_nms|box_embed|patchify|ln_stats|layernorm_kernel|attention_split|gemm_skinny|attention_cls' -c 24 \
  -o /tmp/r2_hbm_final -f python tools/infer_probe.py --images 32 --once --precision f16c8 > gpurun_out/ncu_h.log 2>&1
python tools/ncu_summary.py report /tmp/r2_hbm_final.ncu-rep gpurun_out/r2_hbm_kernels.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2' -c 9 \
  -o /tmp/r2_layer_final -f python tools/infer_probe.py --images 4 --once --precision f16c8 > gpurun_out/ncu_g.log 2>&1
python tools/ncu_summary.py report /tmp/r2_layer_final.ncu-rep gpurun_out/r2_layer_f16c8_ncu.txt
rm -f gpurun_out/r2_launches.csv
du -sh gpurun_out
