cd /root/repo
python scripts/e2e_probe.py > gpurun_out/r2_e2e_probe.txt 2>&1
echo "== synccheck tier 3 (detail)" > gpurun_out/r2_sanitizer2.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 3 python scripts/sanitize_case.py 3 10 2>&1 | grep -v "Host Frame\|^=========$" | head -40 >> gpurun_out/r2_sanitizer2.txt
for tool in memcheck racecheck; do for tier in 2 3; do
  echo "== $tool tier $tier T=70" >> gpurun_out/r2_sanitizer2.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_case.py $tier 70 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|hazard|Invalid" | head -12 >> gpurun_out/r2_sanitizer2.txt
done; done
