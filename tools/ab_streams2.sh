#!/bin/bash
cd "$(dirname "$0")/.."
one() {
  python bench.py --no-cpu-baseline --steps 20 $2 2>gpurun_out/ab_streams.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', '$2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['launch_mode'])"
}
for lim in 36 44 50 66 74 88 100 110 132; do
  DEVIT_SUB_STREAMS=4 DEVIT_SM_LIMIT=$lim one "streams=4 sms=$lim"
done
DEVIT_SUB_STREAMS=4 DEVIT_SM_LIMIT=74 one "streams=4 sms=74" --dense
DEVIT_SUB_STREAMS=4 DEVIT_SM_LIMIT=50 one "streams=4 sms=50" --dense
