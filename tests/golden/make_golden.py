"""Writes the committed regression fixtures tests/golden/*.npz.

Origin of the expected values: the literal Python transliteration of the reference path
(oracle/pilon_oracle.py, every function citing PileUpRegion.scala / PileUp.scala / GenomeRegion.scala), run in the build
container.  The reference itself cannot run there (no JVM / Scala toolchain, and it ships no tests or vectors), so these
are NOT outputs of the reference: they freeze today's restatement so that the C oracle, the packer and the CUDA engine
are all checked against one committed set of numbers, and so that any later drift of the oracle shows up as a diff.
Parity stays "unpinned" in the sense of DESIGN.md section 0.

    python tests/golden/make_golden.py        # rewrites every fixture (deterministic: seeded generators)

Each fixture holds the inputs in the engine's own packed layout (the 13 arrays of pb_batch, per batch) plus the contig,
and the expected per-locus planes, scalars, indel evidence and addRead return values.
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pilon_oracle as po                      # noqa: E402
from pilon_b200.packing import pack_records                # noqa: E402
from tests import helpers as H                             # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BATCH_FIELDS = ("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off", "quals", "bases2",
                "exc_idx", "exc_base", "exc_qual")


def cases():
    """(name, contig, start, stop, [(reads, counts_toward_frag_coverage)], Config)"""
    for seed, cfg in ((0, po.Config()),
                      (3, po.Config(oldIndel=True)),
                      (4, po.Config(minDepth=3.0)),
                      (5, po.Config(fixAmb=True)),
                      (6, po.Config(minQual=7, minMq=2, flank=3, defaultQual=15)),
                      (21, po.Config(flank=0))):
        contig, start, stop, reads = H.random_case(seed)
        yield "random_%02d" % seed, contig, start, stop, H.split_batches(reads, random.Random(seed * 7 + 1)), cfg
    contig, start, stop, reads = H.clean_case(2, n=6000, start=1001, stop=5000, depth=10, n_sites=12)
    yield "clean_02", contig, start, stop, [(reads, True)], po.Config()


def main():
    for name, contig, start, stop, groups, cfg in cases():
        py = H.run_py_oracle(contig, start, stop, groups, cfg)
        packed = [(pack_records(g), f) for g, f in groups]
        res, ins = H.run_c_oracle(contig, start, stop, packed, cfg)        # the two restatements agree before anything is frozen
        H.assert_matches_py(res, ins, py, name)
        out = {"contig": np.frombuffer(contig, np.uint8)}
        for i, (rb, _) in enumerate(packed):
            for f in BATCH_FIELDS:
                out["b%d_%s" % (i, f)] = getattr(rb, f)
        for pname in H.PLANE_NAMES:
            out["plane_" + pname] = np.asarray(py[pname])
        meta = dict(start=start, stop=stop, cfg=dict(vars(cfg)), frag=[bool(f) for _, f in packed],
                    scalars={k: (float(v) if isinstance(v, float) else int(v)) for k, v in py["scalars"].items()},
                    indel_list_len={"%d,%d" % k: int(v) for k, v in py["indel_list_len"].items()},
                    indel_strings={"%d,%d" % k: bytes(v).hex() for k, v in py["indel_strings"].items()},
                    insert_sizes=[int(v) for v in py["insert_sizes"]],
                    per_bam=[[int(x) for x in t] for t in py["per_bam"]])
        out["meta"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-12s %6d loci %5d reads %7d bytes" % (name, stop + 1 - start, sum(rb.n_reads for rb, _ in packed),
                                                     os.path.getsize(path)))


if __name__ == "__main__":
    main()
