for tool in memcheck racecheck synccheck; do
  if [ $tool = memcheck ]; then export SANITIZE_LARGE=1; else unset SANITIZE_LARGE; fi
  /usr/bin/time -f "%e s" timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/r02z_sanitize_$tool.log 2>&1
  echo "compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py (SANITIZE_LARGE=${SANITIZE_LARGE:-0}) -> rc $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run done' gpurun_out/r02z_sanitize_$tool.log | tr '\n' ' ') $(tail -1 gpurun_out/r02z_sanitize_$tool.log)"
done
