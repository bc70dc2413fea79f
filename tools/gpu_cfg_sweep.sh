#!/bin/bash
# chunk / stream configuration sweep of the bench (256 pairs, device-resident value only)
for cfg in "8 1" "8 2" "8 3" "8 4" "4 2" "4 4" "16 2" "16 3" "2 4"; do
  set -- $cfg
  R3D_BENCH_PPC=$1 R3D_BENCH_STREAMS=$2 python bench.py --pairs 256 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ppc $1 streams $2: value %.1f e2e %.1f pairs/s, nn avg launch %.3f ms' % (d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms']))"
done
