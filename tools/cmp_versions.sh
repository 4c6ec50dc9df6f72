#!/bin/bash
# value / roofline / e2e of the default engine on the other BASELINE configs (C2 is bench.py's default)
for wl in C1 C3 C5; do
timeout 600 python bench.py --workload $wl --no-cpu-baseline --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl value %.1f G/s  frac %.3f  ms %.2f e2e %.1f G/s'%(d['value']/1e9, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']/1e9))"
done
