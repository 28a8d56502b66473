set -x
cd $GRAFT_REPO_ROOT
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_job.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log; tail -4 gpurun_out/r2_sanitizer_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_job.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log; tail -4 gpurun_out/r2_sanitizer_racecheck.log
