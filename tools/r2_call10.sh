#!/bin/bash
cd "$(dirname "$0")/.."
one() {
  python bench.py --no-cpu-baseline --no-dense-arm --steps 20 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
}
for m in 0 1; do for st in 0 10000 25000 40000; do
  DEVIT_MLP_STAGGER_MODE=$m DEVIT_MLP_STAGGER=$st one "mode=$m stagger=$st"
done; done
