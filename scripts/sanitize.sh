# compute-sanitizer (memcheck, racecheck, synccheck) over small solves that touch every kernel; one gpurun call:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'      -> gpurun_out/sanitize_<tool>.log + sanitize_summary.txt
CASES=${CASES:-"fused_bs fused_dense staged gjp thomas osc stages"}
TOOLS=${TOOLS:-"memcheck racecheck synccheck"}
: > gpurun_out/sanitize_summary.txt
for tool in $TOOLS; do
  log=gpurun_out/sanitize_${tool}.log
  timeout ${SAN_TIMEOUT:-500} compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py $CASES > $log 2>&1
  echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" >> gpurun_out/sanitize_summary.txt
  grep -E " done$|converged|passed|failed|True|False" $log | sed 's/^/   /' >> gpurun_out/sanitize_summary.txt
done
cat gpurun_out/sanitize_summary.txt
