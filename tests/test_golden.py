"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the literal Python
transliteration of PileUpRegion.scala / PileUp.scala / GenomeRegion.scala): the C oracle on CPU and the CUDA engine
through the C ABI must both reproduce them bit for bit from the stored packed batches."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import pilon_oracle as po
from pilon_b200.packing import ReadBatch
from tests import helpers as H
from tests.golden.make_golden import BATCH_FIELDS

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    batches = [(ReadBatch(**{f: z["b%d_%s" % (i, f)] for f in BATCH_FIELDS}), frag) for i, frag in enumerate(meta["frag"])]
    key = lambda s: tuple(int(x) for x in s.split(","))
    py = {name: z["plane_" + name] for name in H.PLANE_NAMES}
    py["scalars"] = meta["scalars"]
    py["indel_list_len"] = {key(k): v for k, v in meta["indel_list_len"].items()}
    py["indel_strings"] = {key(k): bytes.fromhex(v) for k, v in meta["indel_strings"].items()}
    py["insert_sizes"] = meta["insert_sizes"]
    py["per_bam"] = [tuple(t) for t in meta["per_bam"]]
    return bytes(z["contig"]), meta["start"], meta["stop"], batches, po.Config(**meta["cfg"]), py


def test_fixtures_are_present():
    assert len(GOLDEN) >= 7


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_c_oracle_reproduces_golden(path):
    contig, start, stop, batches, cfg, py = load(path)
    res, ins = H.run_c_oracle(contig, start, stop, batches, cfg)
    H.assert_matches_py(res, ins, py, os.path.basename(path))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_engine_reproduces_golden(path, pileup_kernel):
    from pilon_b200.engine import Engine
    from tests.test_engine_gpu import eng_cfg
    contig, start, stop, batches, cfg, py = load(path)
    e = Engine(0, eng_cfg(cfg))
    try:
        res, ins = e.run_region(contig, start, stop, batches)
        H.assert_matches_py(res, ins, py, "engine vs " + os.path.basename(path))
        rd = [(rb.with_packed_quals().with_base_deltas(contig, start, stop), f) for rb, f in batches]   # compact transports
        res2, ins2 = e.run_region(contig, start, stop, rd)
        H.assert_matches_py(res2, ins2, py, "engine (compact transports) vs " + os.path.basename(path))
    finally:
        e.close()
