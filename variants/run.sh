#!/bin/bash
# tuning harness: bench every library variant in variants/*.so (AGS_B200_LIB override)
for v in "$@"; do
  AGS_B200_LIB=$PWD/variants/$v.so python bench.py --steps 150 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
k=d['kernels']
print('$v', 'step %.1f us' % (d['ms_per_step']*1e3), ' '.join('%s=%.0f' % (n[:11], k[n]['ms']*1e3) for n in k))"
done
