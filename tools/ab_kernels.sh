#!/bin/bash
# A/B of the pileup kernels: tools/ab_kernels.sh WORKLOAD SCALE label:VAR=v,VAR=v ...   -> one line per variant
# (HBM-resident value, ms/step, pileup-kernel ms, roofline fraction, parity against the C oracle)
wl=$1; scale=$2; shift 2
mkdir -p gpurun_out/ab
for spec in "$@"; do
    label=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
    out=gpurun_out/ab/${wl}_${scale}_d${PB_SYNTH_DEPTH:-std}_$label
    env ${envs//,/ } timeout 400 python bench.py --workload $wl --scale $scale --steps 10 --warmup 3 --no-cpu-baseline > $out.json 2> $out.err
    python - "$wl x$scale $label" $out.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "| value %.1f G" % (d["value"] / 1e9), "| ms/step %.3f" % d["ms_per_step"], "| pileup_ms %.3f" % d["roofline"]["pileup_ms_per_step"],
          "| frac %.3f" % d["roofline"]["frac"], "| depth %.0f" % d["config"]["mean_depth"], "| parity", (d.get("parity") or {}).get("ok"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
