# every documented A/B switch of the library must still give parity-green logits
mkdir -p gpurun_out
rm -f gpurun_out/r2_envmatrix.log
for e in "VETO_GEMM_BN256=0" "VETO_QKV_ITEM_LAYOUT=0" "VETO_RESIDUAL_OPERAND=0" "VETO_LN_FUSION=0" "VETO_ATTENTION_SPLIT=0" "VETO_GEMM_CLUSTER4=2" "VETO_LN_STATS_EPILOGUE=1" "VETO_GEMM_GENERIC_EPI=1" "VETO_GEMM_2CTA=0" "VETO_GEMM_PAIRS=20"; do
  r=$(env $e timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "relation_logits or meet_group_heads or chunking" 2>&1 | tail -1)
  echo "$e : $r" | tee -a gpurun_out/r2_envmatrix.log
done
