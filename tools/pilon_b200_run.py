"""`pilon --genome g.fasta --frags f.bam [--jumps j.bam] --fix snps,indels [--changes] [--vcf] [--tracks]` with the B200
engine in place of the reference's pileup path: FASTA + BAM in, .fasta / .changes / .vcf out, chunked and ordered as
GenomeFile.processRegions does (GenomeFile.scala:84-176).  A thin driver over the library's public pieces (bamio, engine,
output); it exists for tools/run_real_pilon.sh and as an end-to-end example -- the Scala driver stays the reference's."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pilon_b200 import bamio, output, synth  # noqa: E402
from pilon_b200.engine import Engine  # noqa: E402


def read_fasta(path):
    name, seq, out = None, [], []
    for ln in open(path, "rb"):
        ln = ln.rstrip(b"\r\n")
        if ln.startswith(b">"):
            if name is not None:
                out.append((name, b"".join(seq)))
            name, seq = ln[1:].split()[0].decode(), []
        else:
            seq.append(ln)
    if name is not None:
        out.append((name, b"".join(seq)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", required=True)
    ap.add_argument("--frags", action="append", default=[])
    ap.add_argument("--jumps", action="append", default=[])
    ap.add_argument("--unpaired", action="append", default=[])
    ap.add_argument("--bam", action="append", default=[])
    ap.add_argument("--changes", action="store_true")
    ap.add_argument("--vcf", action="store_true")
    ap.add_argument("--outdir", default=".")
    ap.add_argument("--output", default="pilon")
    ap.add_argument("--chunksize", type=int, default=synth.CHUNK_SIZE)
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)
    contigs = read_fasta(a.genome)
    bams = [bamio.BamFile(p, t) for t, ps in (("frags", a.frags), ("jumps", a.jumps), ("unpaired", a.unpaired), ("bam", a.bam)) for p in ps]
    eng = Engine(0)
    planes = None if a.vcf else ["flags", "call", "frag_coverage"]
    fasta = open(os.path.join(a.outdir, a.output + ".fasta"), "w")
    changes = open(os.path.join(a.outdir, a.output + ".changes"), "w") if a.changes else None
    vcf = open(os.path.join(a.outdir, a.output + ".vcf"), "w") if a.vcf else None
    if vcf:
        vcf.write(output.vcfHeader(time.strftime("%Y%m%d"), "pilon_b200", " ".join(sys.argv[1:]), "file:" + os.path.abspath(a.genome),
                                   [(n, len(s)) for n, s in contigs]))
    genome_size = sum(len(s) for _, s in contigs)
    for name, seq in contigs:
        outs = []
        for start, stop in synth.chunks_of(len(seq), a.chunksize):
            batches = [(b.process(name, start, stop), b.countsTowardFragCoverage) for b in bams]
            n_ops = sum(int(rb.cigar.shape[0]) for rb, _ in batches)
            res, _ = eng.run_region(seq, start, stop, batches, planes, indels_cap=max(1 << 16, n_ops), bytes_cap=max(1 << 20, 4 * n_ops))
            for b, (_, bc, _) in zip(bams, res.per_bam()):
                b.baseCount += bc                                                   # BamFile.scala:146
            ro = output.RegionOutput(res, seq, name, start, stop)
            print("%s:%d-%d  %s" % (name, start, stop, "; ".join(ro.log())))
            outs.append(ro)
        ch, fa, vc = output.writeContig(name, outs, vcf=bool(vcf), changes=bool(changes))
        fasta.write(fa)
        if changes:
            changes.write("".join(c + "\n" for c in ch))
        if vcf:
            vcf.write(vc)
        for o in outs:
            o.close()
    for ln in output.coverageSummary([(b.bamType, b.baseCount) for b in bams], genome_size):
        print(ln)
    for f in (fasta, changes, vcf):
        if f:
            f.close()
    eng.close()


if __name__ == "__main__":
    main()
