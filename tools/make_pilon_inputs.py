"""Writes a synthetic BASELINE workload as the files Pilon itself reads: <out>/genome.fasta (+ .fai) and one coordinate-sorted
BAM + BAI per library (<out>/frags.bam, <out>/jumps.bam), from the same seeded generator the engine's tests and bench use.

    python tools/make_pilon_inputs.py C1 --scale 0.02 --out /tmp/c1

The files are what tools/run_real_pilon.sh feeds to a Pilon JVM (when one exists) and what `bench.py --from-bam` ingests."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pilon_b200 import bamio, synth  # noqa: E402


def write_inputs(wl, out):
    os.makedirs(out, exist_ok=True)
    names = ["contig%02d" % (i + 1) for i in range(len(wl.contig_lens))]
    contigs = [(names[i], wl.contig_bases(i).tobytes()) for i in range(len(names))]
    bamio.write_fasta(os.path.join(out, "genome.fasta"), contigs)
    refs = [(n, len(s)) for n, s in contigs]
    paths = {}
    for li, libr in enumerate(wl.libraries):
        batches = []
        for ci, n in enumerate(wl.contig_lens):
            sb = synth.SynthBatch(wl.params(ci, libr), 1, n, libr.counts_toward_frag_coverage)     # the whole contig, sorted
            batches.append((ci, sb))
        p = os.path.join(out, "%s.bam" % libr.name)
        bamio.write_bam(p, refs, [(ci, sb.c) for ci, sb in batches], program_line="@PG\tID:pilon_b200_synth\tCL:%s" % wl.name)
        paths[libr.name] = p
        del batches
    return names, paths


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=0.02)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    wl = synth.workload(a.workload, a.scale)
    names, paths = write_inputs(wl, a.out)
    print("wrote %d contigs to %s/genome.fasta; BAMs: %s" % (len(names), a.out, ", ".join("%s=%s" % kv for kv in paths.items())))
