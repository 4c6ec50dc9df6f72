for wl in C1 C3 C5; do for v in 5 7; do
PB_PILEUP=$v timeout 600 python bench.py --workload $wl --no-cpu-baseline --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl v$v value %.1f G/s  frac %.3f  ms %.2f e2e %.1f'%(d['value']/1e9, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']/1e9))"
done; done
