#!/bin/bash
# The tuning switches that exist but have not been measured (or only at N = 2), one bench line each.
#   1 GPU :  gpurun --timeout 900 -- 'bash tools/variants_probe.sh 1 > gpurun_out/variants1.log 2>&1'
#   N GPUs:  gpurun --gpus N --timeout 900 -- 'bash tools/variants_probe.sh N > gpurun_out/variantsN.log 2>&1'
# ZB_MM_CFG     mirror_merge_kernel shape: 0 = 256 threads / 1024 groups (default), 1 = 512 / 2048, 2 = 256 / 2048
# ZB_P2P_RESERVE  1 = the exchange without a count matrix (runs reserved by system-scope atomics in the owner's buffer)
# ZB_ROUTE_PER  keys per thread of a routing tile in reserve mode: 8 (default) or 16 (half as many reservations)
set -u
cd "$(dirname "$0")/.."
N=${1:-1}
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); st=d['roofline']['stage_ms_per_step']
        print('%-34s %.2f Gbases/s  %.3f ms/step  mirror_buckets %.3f  route_p2p %s' % (sys.argv[1], d['value'], d['ms_per_step'], st.get('mirror_buckets', 0), st.get('route_p2p', '-')))
" "$1"; }
if [ "$N" -eq 1 ]; then
    for cfg in 0 1 2; do
        ZB_MM_CFG=$cfg python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pairs 2>/dev/null | line "ZB_MM_CFG=$cfg"
    done
else
    for v in "ZB_P2P_RESERVE=0" "ZB_P2P_RESERVE=1 ZB_ROUTE_PER=8" "ZB_P2P_RESERVE=1 ZB_ROUTE_PER=16"; do
        env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
            bench.py --gpus "$N" --steps 8 --warmup 3 --no-cpu-baseline --no-pairs 2>/dev/null | line "N=$N $v"
    done
fi
