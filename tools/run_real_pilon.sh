#!/usr/bin/env bash
# Diffs the outputs of the REAL Pilon (a JVM + pilon jar, neither of which exists in the build image) against this
# engine's on identical synthetic inputs: the only route by which parity can stop being "unpinned" (DESIGN.md section 0).
#
#   tools/run_real_pilon.sh [WORKLOAD] [SCALE]        e.g.  tools/run_real_pilon.sh C1 0.02
#
# Looks for `java` on PATH and a pilon jar under baseline/_ref/ (the place .gitignore reserves for a driver-installed
# reference).  Without them it says so and exits 0: nothing can be compared.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
WL="${1:-C1}"; SCALE="${2:-0.02}"
JAR="$(ls "$ROOT"/baseline/_ref/*.jar 2>/dev/null | head -1 || true)"
if ! command -v java >/dev/null 2>&1 || [ -z "$JAR" ]; then
  echo "run_real_pilon: no JVM and/or no baseline/_ref/*.jar on this machine -- the reference cannot run; parity stays unpinned"
  exit 0
fi
OUT="$(mktemp -d)"
python "$ROOT/tools/make_pilon_inputs.py" "$WL" --scale "$SCALE" --out "$OUT/in"
ARGS=(--genome "$OUT/in/genome.fasta" --fix snps,indels --changes --vcf --outdir "$OUT/ref" --output pilon)
[ -f "$OUT/in/frags.bam" ] && ARGS+=(--frags "$OUT/in/frags.bam")
[ -f "$OUT/in/jumps.bam" ] && ARGS+=(--jumps "$OUT/in/jumps.bam")
java -Xmx32G -jar "$JAR" "${ARGS[@]}" > "$OUT/ref.log"
python "$ROOT/tools/pilon_b200_run.py" --genome "$OUT/in/genome.fasta" $( [ -f "$OUT/in/frags.bam" ] && echo --frags "$OUT/in/frags.bam" ) \
       $( [ -f "$OUT/in/jumps.bam" ] && echo --jumps "$OUT/in/jumps.bam" ) --changes --vcf --outdir "$OUT/b200" --output pilon
rc=0
for ext in fasta changes; do
  if cmp -s "$OUT/ref/pilon.$ext" "$OUT/b200/pilon.$ext"; then echo "pilon.$ext: identical"; else echo "pilon.$ext: DIFFERENT"; rc=1; fi
done
# the VCF header carries the date, the version string and the command line: compare the records only
if diff <(grep -v '^##' "$OUT/ref/pilon.vcf") <(grep -v '^##' "$OUT/b200/pilon.vcf") >/dev/null; then echo "pilon.vcf records: identical"; else echo "pilon.vcf records: DIFFERENT"; rc=1; fi
echo "outputs kept under $OUT"
exit $rc
