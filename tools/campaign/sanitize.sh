for tool in memcheck racecheck synccheck; do
  if [ $tool = memcheck ]; then export SANITIZE_LARGE=1; else unset SANITIZE_LARGE; fi
  t0=$SECONDS
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r02z_sanitize_$tool.log 2>&1
  rc=$?
  echo "compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py (SANITIZE_LARGE=${SANITIZE_LARGE:-0}) -> rc $rc, $((SECONDS - t0)) s : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run done' gpurun_out/r02z_sanitize_$tool.log | tr '\n' ' ')"
done
