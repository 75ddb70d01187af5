mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( time timeout 150 python scripts/sanitize_trunk.py ) > gpurun_out/r5a_trunk_plain.log 2>&1; echo "plain trunk rc=$?"; tail -4 gpurun_out/r5a_trunk_plain.log | cut -c1-300
( time timeout 200 compute-sanitizer --tool memcheck python scripts/sanitize_capture.py ) > gpurun_out/r5a_memcheck_capture.log 2>&1; echo "memcheck capture rc=$?"; tail -6 gpurun_out/r5a_memcheck_capture.log | cut -c1-300
( time timeout 240 compute-sanitizer --tool memcheck python scripts/sanitize_trunk.py ) > gpurun_out/r5a_memcheck_trunk.log 2>&1; echo "memcheck trunk rc=$?"; tail -6 gpurun_out/r5a_memcheck_trunk.log | cut -c1-300
( time timeout 150 compute-sanitizer --tool racecheck python scripts/sanitize_capture.py ) > gpurun_out/r5a_racecheck_capture.log 2>&1; echo "racecheck capture rc=$?"; tail -6 gpurun_out/r5a_racecheck_capture.log | cut -c1-300
