#!/bin/bash
# compute-sanitizer over the small end-to-end invocation of __graft_entry__.smoke() (kmerize + count, trim, merge, pair and
# all-pairs cardinalities, codec64 streams -- ~350 launches on a 200 kbp genome), SURVEY.md section 5 "race detection /
# sanitizers".  Run on a GPU box from the repository root:
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1; tail -30 gpurun_out/sanitize.log'
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the chained scans of sort.cu / parse.cu and
# the hash tables of segsort.cu / allpairs.cu are where they would be); synccheck: divergent barriers; initcheck: reads of
# uninitialised global memory.  Each tool's summary line is echoed at the end.  NOT run in round 1 (the GPU budget of the round was
# spent before this script existed) -- see DESIGN.md section 7.
set -u
cd "$(dirname "$0")/.."
rc=0
for tool in memcheck racecheck synccheck initcheck; do
    echo "==== compute-sanitizer --tool $tool"
    timeout 1200 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 3 \
        python -c "import __graft_entry__ as g; g.smoke()" > "/tmp/sanitize_$tool.log" 2>&1
    code=$?
    tail -n 8 "/tmp/sanitize_$tool.log"
    echo "==== $tool exit code $code"
    [ "$code" -ne 0 ] && rc=1
done
exit $rc
