#!/bin/bash
# A/B: how many CUDA streams a rank's sub-models are spread over (DEVIT_SUB_STREAMS), optionally
# with every persistent grid sized for half the chip (DEVIT_SM_LIMIT=74).
#   gpurun --timeout 600 -- 'bash tools/ab_streams.sh > gpurun_out/ab_streams.txt 2>&1'
cd "$(dirname "$0")/.."
one() {
  python bench.py --no-cpu-baseline --steps 20 $2 2>gpurun_out/ab_streams.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', '$2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['launch_mode'])"
}
for st in 1 2 4; do
  DEVIT_SUB_STREAMS=$st one "streams=$st"
done
DEVIT_SUB_STREAMS=2 DEVIT_SM_LIMIT=74 one "streams=2 sms=74"
DEVIT_SUB_STREAMS=4 DEVIT_SM_LIMIT=74 one "streams=4 sms=74"
DEVIT_SUB_STREAMS=1 one "streams=1" --dense
DEVIT_SUB_STREAMS=2 one "streams=2" --dense
