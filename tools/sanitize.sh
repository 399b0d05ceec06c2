#!/bin/bash
# compute-sanitizer over the small end-to-end invocation of __graft_entry__.smoke() (kmerize + count, trim, merge, pair and
# all-pairs cardinalities, codec64 streams -- ~350 launches on a 200 kbp genome), SURVEY.md section 5 "race detection /
# sanitizers".  Run on a GPU box from the repository root:
#     gpurun --timeout 1800 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1; tail -30 gpurun_out/sanitize.log'
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the chained scans of sort.cu / parse.cu and
# the hash tables of segsort.cu / allpairs.cu are where they would be); synccheck: divergent barriers; initcheck: reads of
# uninitialised global memory.  The full log of every tool is kept under gpurun_out/ (summaries are copied to profiles/).
# Optional argument: a python statement to run instead of smoke() (e.g. a 2-rank check is run by tools/sanitize_mgpu.sh).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${ZB_SAN_TOOLS:-"memcheck synccheck initcheck racecheck"}
STMT=${1:-"import __graft_entry__ as g; g.smoke()"}
TAG=${ZB_SAN_TAG:-smoke}
rc=0
for tool in $TOOLS; do
    echo "==== compute-sanitizer --tool $tool"
    t0=$(date +%s)
    timeout ${ZB_SAN_TIMEOUT:-420} compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 3 \
        python -c "$STMT" > "gpurun_out/sanitize_${TAG}_$tool.log" 2>&1
    code=$?
    tail -n 8 "gpurun_out/sanitize_${TAG}_$tool.log"
    echo "==== $tool exit code $code after $(( $(date +%s) - t0 )) s"
    [ "$code" -ne 0 ] && rc=1
done
exit $rc
