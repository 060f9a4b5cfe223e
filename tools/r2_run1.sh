# round 2, first GPU call: sanity, chunk-size sweep of the inference step, ncu captures of the HBM kernels (before the rebuild)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 600 python tools/infer_probe.py --images 32 --chunks 0,499,997,3988,7976 --steps 2 --warmup 1 > gpurun_out/r2_chunk_sweep.jsonl 2> gpurun_out/r2_chunk_sweep.err
cat gpurun_out/r2_chunk_sweep.jsonl; tail -3 gpurun_out/r2_chunk_sweep.err
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'roi_gather|pairs_|tokens_kernel|postprocess_kernel|obj_nms|box_embed|patchify' -c 24 \
  -o gpurun_out/r2_hbm_before -f python tools/infer_probe.py --images 4 --once > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'layernorm_kernel|attention_mma2|gemm_skinny|attention_cls' -c 8 \
  -o gpurun_out/r2_rows_before -f python tools/infer_probe.py --images 4 --once > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
ls -la gpurun_out/*.ncu-rep
