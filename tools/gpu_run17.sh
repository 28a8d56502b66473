set -x
cd $GRAFT_REPO_ROOT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_job.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log; tail -6 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_job.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log; tail -6 gpurun_out/r2_sanitizer_racecheck.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
