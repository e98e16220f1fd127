cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mstdp_conv.py -x -q -k "hebbian or conv_postpre or conv_wdep or conv_hebbian or conv_mstdp or lif_wdep" 2>&1 | tail -4 > gpurun_out/r2_tests_b.txt
for tool in memcheck racecheck synccheck; do for tier in 2 3 1; do
  echo "== $tool tier $tier" >> gpurun_out/r2_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_case.py $tier 10 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|hazard|Invalid" | head -12 >> gpurun_out/r2_sanitizer.txt
done; done
