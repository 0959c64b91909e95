#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over the smoke test and the small parity cases
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/${TAG}_sanitizer_racecheck.log
grep -c "ERROR SUMMARY" gpurun_out/${TAG}_sanitizer_*.log; grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitizer_*.log
