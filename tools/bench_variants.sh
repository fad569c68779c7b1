#!/bin/bash
# run under gpurun: count-kernel time of every variant built by tools/variants.sh
for f in neurokmer_b200/build/variants/lib_*.so; do
  NEUROKMER_LIB=$PWD/$f python bench.py --steps 30 --no-cpu --no-e2e --no-parity 2>/tmp/err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read());print('$f', 'count_ms', round(d['phases_ms']['count_ms'],4), 'step_ms', round(d['ms_per_step'],4))
except Exception as e:
    print('$f', 'FAILED', open('/tmp/err.log').read()[-300:].replace(chr(10),' | '))"
done
