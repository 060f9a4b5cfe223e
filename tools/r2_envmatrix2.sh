for e in "VETO_GEMM_2CTA=0" "VETO_NOTHING=1"; do
  r=$(env $e timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "relation_logits or meet_group_heads or chunking or gemm_tcgen05" 2>&1 | tail -1)
  echo "$e : $r"
done
