set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 3 --log-file gpurun_out/r2_memcheck_%p.log python -m pytest tests -q -m gpu -x 2>&1 | tail -5
echo "exit $?"
grep -h "ERROR SUMMARY" gpurun_out/r2_memcheck_*.log | sort | uniq -c
grep -l "Invalid\|out of bounds\|misaligned" gpurun_out/r2_memcheck_*.log | head
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --log-file gpurun_out/r2_racecheck.log python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "(relation_logits and f16c8 and cfg1_predcls_vg) or (relation_logits and bf16x3 and ragged)" 2>&1 | tail -3
tail -5 gpurun_out/r2_racecheck.log
ls gpurun_out/r2_memcheck_*.log | wc -l; du -sh gpurun_out
