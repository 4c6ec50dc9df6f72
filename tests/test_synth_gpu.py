"""GPU parity on the synthetic BASELINE workloads (small scale, against the C oracle) and
size-independent properties at a BASELINE-sized region."""
import numpy as np
import pytest

from pilon_b200 import _capi as capi
from pilon_b200 import synth
from pilon_b200.engine import Engine
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _kernel(pileup_kernel):
    return pileup_kernel


@pytest.fixture(scope="module")
def engine(pileup_kernel):
    e = Engine(0)
    yield e
    e.close()


def _check_region(engine, wl, ci, a, b):
    contig = wl.contig_bases(ci).tobytes()
    batches = [(sb.as_read_batch(), sb.frag) for sb in wl.region_batches(ci, a, b)]
    res, ins = engine.run_region(contig, a, b, batches, indels_cap=1 << 18, bytes_cap=1 << 22)
    ref, ins_ref = H.run_c_oracle(contig, a, b, batches, indels_cap=1 << 18, bytes_cap=1 << 22)
    H.assert_results_equal(res, ref, "%s contig %d %d-%d" % (wl.name, ci, a, b))
    for x, y in zip(ins, ins_ref):
        assert np.array_equal(x, y)
    return res


@pytest.mark.parametrize("name,scale", [("C1", 0.02), ("C2", 0.003), ("C3", 0.002), ("C5", 0.2)])
def test_baseline_configs_small_scale(engine, name, scale):
    wl = synth.workload(name, scale)
    seen = 0
    for ci, a, b in wl.regions()[:4]:
        res = _check_region(engine, wl, ci, a, b)
        seen += int((res["flags"] & (capi.PB_FL_CHANGED | capi.PB_FL_AMBIGUOUS)).astype(bool).sum())
    assert seen > 0, "planted variants must produce calls"


def test_chunked_contig_with_halo_reads(engine):
    # one contig cut into 4 chunks: every chunk sees reads that start in its neighbours (+-10 kb halo)
    wl = synth.workload("C3", 0.002)
    wl.chunk_size = 40_000
    regs = wl.regions()
    assert len(regs) == 4 and regs[1][1] > 1
    for ci, a, b in regs:
        _check_region(engine, wl, ci, a, b)


def test_full_size_region_properties(engine):
    """BASELINE-sized unit (C1: 5 Mb at 100x = 0.5 G aligned bases): too big for the oracle in a test,
    so check properties that do not depend on size."""
    wl = synth.workload("C1")
    ci, a, b = wl.regions()[0]
    contig = wl.contig_bases(ci)
    sb = wl.region_batches(ci, a, b)[0]
    planes = ["base_count4", "deletions", "coverage_arr", "bad_pair", "frag_coverage", "phys_cov", "flags", "call", "mq_sum", "q_sum"]
    res, _ = engine.run_region(contig, a, b, [(sb, True)], planes=planes, indels_cap=1 << 20, bytes_cap=1 << 24)
    cnt = res["base_count4"].astype(np.int64).sum(axis=1)
    dele = res["deletions"].astype(np.int64)
    # depth = count + deletions (PileUp.scala:44) is what pass 1 stores as coverage
    assert np.array_equal(res["coverage_arr"].astype(np.int64), cnt + dele)
    assert res.c.read_count == sb.n_reads and res.c.aligned_bases == sb.aligned_bases
    # region baseCount counts every trusted base of valid reads, countable or not (PileUpRegion.scala:41-43)
    assert cnt.sum() <= res.c.base_count <= sb.aligned_bases
    assert res.c.coverage == (res.c.base_count + (b - a + 1) // 2) // (b - a + 1)
    # one BAM, not "jumps": fragCoverage is the depth before the deletion spill
    fl = res["flags"]
    not_deleted = (fl & capi.PB_FL_DELETED) == 0
    assert np.array_equal(res["frag_coverage"][not_deleted], res["coverage_arr"][not_deleted])
    # physical coverage is a prefix sum of +1/-1 per fragment: non-negative, bounded by the fragment count
    assert res["phys_cov"].min() >= 0 and res["phys_cov"].max() <= sb.n_reads
    # split invariance: the same reads in three batches give the same result
    rb = sb.as_read_batch()
    n = rb.n_reads
    parts = [(rb.slice_reads(0, n // 3), True), (rb.slice_reads(n // 3, 2 * n // 3), True), (rb.slice_reads(2 * n // 3, n), True)]
    res3, _ = engine.run_region(contig, a, b, parts, planes=planes, indels_cap=1 << 20, bytes_cap=1 << 24)
    for p in planes:
        assert np.array_equal(res[p], res3[p]), p
    assert (res3.c.base_count, res3.c.read_count, res3.c.min_depth) == (res.c.base_count, res.c.read_count, res.c.min_depth)
    # the planted variants are found: ~1 SNP per kb
    snps = int((((fl & capi.PB_FL_CHANGED) != 0) & (((fl >> capi.PB_FL_KIND_SHIFT) & 3) == capi.PB_KIND_SNP)).sum())
    assert 0.8 * 5000 <= snps <= 1.2 * 5000
