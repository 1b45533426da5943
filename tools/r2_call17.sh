#!/bin/bash
cd "$(dirname "$0")/.."
one() {
  python bench.py --no-cpu-baseline --no-dense-arm --steps 20 $2 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', '$2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
}
for sh in 2 1; do for st in 4 2; do
  DEVIT_SM_SHARE=$sh DEVIT_SUB_STREAMS=$st one "share=$sh streams=$st"
done; done
DEVIT_SM_SHARE=2 DEVIT_SUB_STREAMS=4 one "share=2 streams=4" --dense
DEVIT_SM_SHARE=1 DEVIT_SUB_STREAMS=4 one "share=1 streams=4" --dense
DEVIT_SM_SHARE=3 DEVIT_SUB_STREAMS=4 one "share=3 streams=4"
DEVIT_SUB_STREAMS=1 one "single chain"
