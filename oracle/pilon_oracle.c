/*
 * ORACLE (test infrastructure, NOT product code) -- C restatement of Pilon's pileup + BaseCall
 * hot path, operating on the same packed read batches (include/pilon_b200.h) as the CUDA engine.
 *
 *   *** parity unpinned ***  The reference ships no tests / golden vectors and cannot run in the
 *   build container (no JVM).  This file restates the Scala read-by-read, exactly as the JVM
 *   executes it (scatter per read, sequential pass 1); it is pinned by the hand-derived KATs of
 *   SURVEY.md 8(c) and cross-checked against the literal Python restatement
 *   (oracle/pilon_oracle.py) by tests/test_oracle_cross.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under pilon_b200/ does.
 *
 * Citations "File.scala:a-b" are relative to
 * /root/reference/src/main/scala/org/broadinstitute/pilon/ .
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../include/pilon_b200.h"

/* ---- JVM arithmetic ------------------------------------------------------------------- */
static inline int32_t wrap32(int64_t x) { return (int32_t)(uint32_t)(uint64_t)x; }
/* Utils.scala:23-26; C99 '/' truncates toward zero like the JVM */
static inline int64_t roundDivL(int64_t n, int64_t d) { return d > 0 ? (int64_t)((uint64_t)n + (uint64_t)(d / 2)) / d : 0; }
static inline int32_t roundDivI(int32_t n, int32_t d) { return d > 0 ? wrap32((int64_t)n + d / 2) / d : 0; }
static inline int32_t pctI(int32_t n, int32_t d) { return roundDivI(wrap32(100LL * n), d); }

typedef struct {
    int32_t loc;      /* region index */
    int32_t kind;     /* 1 ins, 2 del */
    int32_t len;
    int32_t pad;
    uint8_t* bytes;   /* owned */
} po_event;

typedef struct po_region {
    pb_config cfg;
    const uint8_t* contig;     /* not owned */
    int64_t contig_len;
    int32_t start, stop;
    int64_t size;
    /* PileUp fields, SoA (PileUp.scala:26-39) */
    int64_t *bc, *qs;          /* [size*4] */
    int32_t *mqSum, *qSum, *physCov, *insertSize, *badPair, *deletions, *delQual, *insertions, *insQual, *clips;
    int32_t *fragCov;
    int32_t *covBefore;
    int64_t baseCount;         /* PileUpRegion.scala:32 */
    int32_t readCount;         /* :33 */
    int32_t physCovStart, insertSizeStart;   /* :59-60 */
    int32_t unknown_ops, dropped_oob;
    int64_t aligned_bases;
    po_event* ev; int64_t n_ev, cap_ev;
    /* per-BAM deltas taken around the read loop (BamFile.scala:120-122,142-146) */
    int32_t bam_reads[256]; int64_t bam_bases[256], bam_cov[256]; int n_bams;
} po_region;

static void* xcalloc(size_t n, size_t s) { void* p = calloc(n ? n : 1, s); if (!p) abort(); return p; }

po_region* po_region_new(const pb_config* cfg, const uint8_t* contig, int64_t contig_len,
                         int32_t start, int32_t stop) {
    po_region* r = (po_region*)xcalloc(1, sizeof(po_region));
    r->cfg = *cfg; r->contig = contig; r->contig_len = contig_len;
    r->start = start; r->stop = stop; r->size = (int64_t)stop + 1 - start;   /* Region.scala:27 */
    size_t n = (size_t)r->size;
    r->bc = xcalloc(n * 4, 8); r->qs = xcalloc(n * 4, 8);
    r->mqSum = xcalloc(n, 4); r->qSum = xcalloc(n, 4); r->physCov = xcalloc(n, 4);
    r->insertSize = xcalloc(n, 4); r->badPair = xcalloc(n, 4); r->deletions = xcalloc(n, 4);
    r->delQual = xcalloc(n, 4); r->insertions = xcalloc(n, 4); r->insQual = xcalloc(n, 4);
    r->clips = xcalloc(n, 4); r->fragCov = xcalloc(n, 4); r->covBefore = xcalloc(n, 4);
    return r;
}

void po_region_free(po_region* r) {
    if (!r) return;
    for (int64_t i = 0; i < r->n_ev; i++) free(r->ev[i].bytes);
    free(r->ev);
    free(r->bc); free(r->qs); free(r->mqSum); free(r->qSum); free(r->physCov); free(r->insertSize);
    free(r->badPair); free(r->deletions); free(r->delQual); free(r->insertions); free(r->insQual);
    free(r->clips); free(r->fragCov); free(r->covBefore);
    free(r);
}

static inline int inRegion(const po_region* r, int64_t locus) { return locus >= r->start && locus <= r->stop; }  /* Region.scala:23 */
static inline int64_t depthAt(const po_region* r, int64_t i) {   /* PileUp.scala:44 */
    const int64_t* b = r->bc + 4 * i;
    return b[0] + b[1] + b[2] + b[3] + r->deletions[i];
}

/* PileUp.add, PileUp.scala:75-84; base is the ASCII read byte, qual the signed JVM byte */
static inline void pileup_add(po_region* r, int64_t i, int base, int qual, int mq) {
    int bi = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : base == 'T' ? 3 : -1;   /* :46-52 */
    if (bi >= 0 && qual >= r->cfg.min_qual) {
        int32_t mq1 = mq + 1;
        r->bc[4 * i + bi] += 1;
        r->qs[4 * i + bi] += (int64_t)wrap32((int64_t)qual * mq1);
        r->mqSum[i] = wrap32((int64_t)r->mqSum[i] + mq1);
        r->qSum[i] = wrap32((int64_t)r->qSum[i] + qual);
    }
}

/* PileUpRegion.add, PileUpRegion.scala:38-48 */
static inline void region_add(po_region* r, int64_t locus, int base, int qual, int mq, int pair) {
    if (inRegion(r, locus)) {
        int64_t i = locus - r->start;
        if (pair) { pileup_add(r, i, base, qual, mq); r->baseCount += 1; }
        else r->badPair[i] = wrap32((int64_t)r->badPair[i] + 1);
    }
}

static void push_event(po_region* r, int32_t loc, int kind, const uint8_t* bytes, int32_t len) {
    if (r->n_ev == r->cap_ev) {
        r->cap_ev = r->cap_ev ? r->cap_ev * 2 : 1024;
        r->ev = (po_event*)realloc(r->ev, (size_t)r->cap_ev * sizeof(po_event));
        if (!r->ev) abort();
    }
    po_event* e = &r->ev[r->n_ev++];
    e->loc = loc; e->kind = kind; e->len = len; e->pad = 0;
    e->bytes = (uint8_t*)malloc(len > 0 ? (size_t)len : 1);
    memcpy(e->bytes, bytes, (size_t)len);
}

/* PileUpRegion.physCovIncr, PileUpRegion.scala:62-88 */
static int32_t physCovIncr(po_region* r, int32_t aStart, int32_t aEnd, int32_t iSize, int paired, int valid) {
    if (!valid || (paired && iSize <= 0)) return 0;
    int64_t s, e;
    if (!paired) { s = aStart < aEnd ? aStart : aEnd; e = aStart > aEnd ? aStart : aEnd; }
    else if (iSize > 0) { s = aStart; e = (int64_t)aStart + iSize; }
    else { s = (int64_t)aEnd + 1 + iSize; e = (int64_t)aEnd + 1; }
    int32_t ins = wrap32(e - s);
    s = wrap32(s); e = wrap32(e);
    if (inRegion(r, s)) {
        int64_t i = s - r->start;
        r->physCov[i] = wrap32((int64_t)r->physCov[i] + 1);
        r->insertSize[i] = wrap32((int64_t)r->insertSize[i] + ins);
    } else if (s < r->start && !(e < r->start)) {
        r->physCovStart = wrap32((int64_t)r->physCovStart + 1);
        r->insertSizeStart = wrap32((int64_t)r->insertSizeStart + ins);
    }
    if (inRegion(r, e)) {
        int64_t i = e - r->start;
        r->physCov[i] = wrap32((int64_t)r->physCov[i] - 1);
        r->insertSize[i] = wrap32((int64_t)r->insertSize[i] - ins);
    }
    return ins;
}

/* Unpack one read into ASCII bases + raw quality bytes (what htsjdk hands to addRead). */
static void unpack_read(const pb_batch* b, int64_t rd, uint8_t* bases, uint8_t* quals, int64_t* exc_cursor) {
    int32_t len = b->read_len[rd];
    uint32_t off = b->seq_off[rd];
    int64_t k = *exc_cursor;
    while (k < b->n_exc && b->exc_idx[k] < off) k++;
    for (int32_t j = 0; j < len; j++) {
        uint32_t i = off + (uint32_t)j;
        uint8_t q = b->quals[i];
        if (q & 0x80) {
            while (k < b->n_exc && b->exc_idx[k] < i) k++;
            /* contract: every bit-7 base has an exception entry */
            bases[j] = b->exc_base[k]; quals[j] = b->exc_qual[k];
        } else {
            bases[j] = (uint8_t)"ACGT"[(b->bases2[i >> 2] >> (2 * (i & 3))) & 3];
            quals[j] = q;
        }
    }
    *exc_cursor = k;
}

/* PileUpRegion.scala:120-134: both helpers index refBases (the whole contig, 0-based) with whatever they are handed -- a
 * locus for insertions (:160), a REGION index for deletions and aligned bases (:180-181,190).  Outside the contig the JVM
 * would throw; here such an index simply matches nothing. */
static int refAt(const po_region* r, int64_t i0) { return (i0 >= 0 && i0 < r->contig_len) ? r->contig[i0] : 0; }
static int homoRun(const po_region* r, int64_t i0) {
    if (i0 < 0 || i0 >= r->contig_len) return 0;
    int b = r->contig[i0];
    for (int64_t i = i0 + 1; i < r->contig_len; i++) if (r->contig[i] != b) return (int)(i - i0 > 1000000 ? 1000000 : i - i0);
    return (int)(r->contig_len - i0 > 1000000 ? 1000000 : r->contig_len - i0);
}
static int nanoporeExclude(const po_region* r, int64_t i0) {
    return inRegion(r, r->start + i0 - 2) && inRegion(r, r->start + i0 + 2) &&
           refAt(r, i0 - 2) == 'C' && refAt(r, i0 - 1) == 'C' && refAt(r, i0 + 1) == 'G' && refAt(r, i0 + 2) == 'G';
}

/* PileUpRegion.addRead, PileUpRegion.scala:102-220; longRead = BamFile.longReadType (0, 1 nanopore, 2 pacbio) */
static int32_t addRead(po_region* r, const pb_batch* b, int64_t rd, const uint8_t* bases, const uint8_t* qraw, int longRead) {
    const pb_config* cfg = &r->cfg;
    const uint8_t* ref = r->contig;
    int32_t length = b->read_len[rd];
    int mq = b->mapq[rd];
    uint8_t fl = b->flags[rd];
    int paired = (fl & PB_F_PAIRED) != 0;
    int valid = (mq >= cfg->min_mq) && (!paired || ((fl & PB_F_PROPER) && (fl & PB_F_MATE_SAME_REF)));   /* :107 */
    int32_t insert = b->tlen[rd];
    int32_t aStart = b->pos[rd];
    int hasq = (fl & PB_F_HAS_QUALS) != 0;
    int32_t flank = cfg->flank;
    const uint32_t* cig = b->cigar + b->cigar_off[rd];
    int32_t ncig = (int32_t)(b->cigar_off[rd + 1] - b->cigar_off[rd]);
#define QUAL(o) (hasq ? (int)(int8_t)qraw[o] : (int)(int8_t)cfg->default_qual)          /* :114-115 */
#define TRUSTED(o) ((o) >= flank && length - flank > (o))                             /* :118 */
    int64_t clipped = 0, reflen = 0;
    for (int32_t k = 0; k < ncig; k++) {
        int op = cig[k] & 15; int64_t len = cig[k] >> 4;
        if (op == 4) clipped += len;                                                  /* :139 */
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += len;
        if (op == 0 || op == 7 || op == 8) r->aligned_bases += len;
    }
    int32_t aEnd = (fl & PB_F_UNMAPPED) ? 0 : wrap32((int64_t)aStart + reflen - 1);   /* getAlignmentEnd */
    int32_t adjMq = roundDivI(wrap32((int64_t)mq * (length - clipped)), length);      /* :141 */
    int32_t indelMq = longRead > 0 ? (adjMq < 8 ? adjMq : 8) : adjMq;                 /* :142 */
    int64_t readOffset = 0, refOffset = 0;
    for (int32_t k = 0; k < ncig; k++) {
        int op = cig[k] & 15; int64_t len = cig[k] >> 4;
        int64_t locus = (int64_t)aStart + refOffset;                                  /* :148 */
        switch (op) {
        case 1: { /* I, :150-162 */
            int64_t iloc = locus;
            if (valid && TRUSTED(readOffset) && inRegion(r, iloc)) {
                uint8_t* ins = (uint8_t*)malloc((size_t)len);
                memcpy(ins, bases + readOffset, (size_t)len);
                while (iloc > 1 && ref[iloc - 2] == ins[len - 1]) {
                    iloc -= 1;
                    uint8_t last = ins[len - 1];
                    memmove(ins + 1, ins, (size_t)(len - 1));
                    ins[0] = last;
                    if (iloc < r->start) break;   /* JVM would crash at pileups(index(iloc)) below */
                }
                if (iloc < r->start) { r->dropped_oob++; free(ins); break; }
                if (longRead > 0 && homoRun(r, iloc) >= 4) { free(ins); break; }         /* :160 (the locus is used as an index) */
                int64_t i = iloc - r->start;
                /* PileUp.addInsertion, PileUp.scala:98-105 */
                r->insQual[i] = wrap32((int64_t)r->insQual[i] + indelMq + 1);
                r->qSum[i] = wrap32((int64_t)r->qSum[i] + QUAL(readOffset));
                r->insertions[i] = wrap32((int64_t)r->insertions[i] + 1);
                push_event(r, (int32_t)i, 1, ins, (int32_t)len);
                free(ins);
            }
            break; }
        case 2: { /* D, :163-183 */
            int64_t dloc = locus, rloc = readOffset;
            if (valid && TRUSTED(readOffset) && inRegion(r, dloc) && inRegion(r, dloc + len - 1)) {
                /* first find where the shift ends: if it leaves the region the JVM dies at :182,
                 * so the op contributes nothing in our defined behaviour */
                int64_t d2 = dloc, r2 = rloc;
                while (d2 > 1 && r2 > 0 && ref[d2 - 2] == ref[d2 + len - 2]) { d2--; r2--; if (d2 < r->start) break; }
                if (d2 < r->start) { r->dropped_oob++; break; }
                while (dloc > 1 && rloc > 0 && ref[dloc - 2] == ref[dloc + len - 2]) {
                    dloc -= 1; rloc -= 1;
                    int base = bases[rloc]; int qual = QUAL(rloc);
                    if (TRUSTED(rloc) && inRegion(r, dloc)) {
                        /* remove(dloc, ..., valid=true) is a no-op, :50-58 */
                        if (inRegion(r, dloc + len)) region_add(r, dloc + len, base, qual, adjMq, valid);
                    }
                }
                int64_t i = dloc - r->start;
                if (longRead > 0 && (homoRun(r, i) >= 4 || (longRead == 1 && nanoporeExclude(r, i)))) break;   /* :180-181 */
                /* PileUp.addDeletion, PileUp.scala:107-114 */
                r->mqSum[i] = wrap32((int64_t)r->mqSum[i] + indelMq + 1);
                r->delQual[i] = wrap32((int64_t)r->delQual[i] + indelMq + 1);
                r->qSum[i] = wrap32((int64_t)r->qSum[i] + QUAL(readOffset));
                r->deletions[i] = wrap32((int64_t)r->deletions[i] + 1);
                push_event(r, (int32_t)i, 2, ref + dloc - 1, (int32_t)len);             /* refBases.slice(dloc-1, dloc+len-1) */
            }
            break; }
        case 0: case 7: case 8: /* M = X, :184-193 */
            for (int64_t i = 0; i < len; i++) {
                int64_t rOff = readOffset + i;
                if (TRUSTED(rOff)) {
                    int qual = (longRead == 1 && nanoporeExclude(r, locus + i - r->start)) ? 0 : QUAL(rOff);   /* :190 */
                    region_add(r, locus + i, bases[rOff], qual, adjMq, valid);
                }
            }
            break;
        case 4: { /* S, :194-206 */
            int64_t clipStart = readOffset == 0 ? locus - len : locus;
            int64_t clipEnd = clipStart + len - 1;
            if (inRegion(r, clipStart)) { int64_t i = clipStart - r->start; r->clips[i] = wrap32((int64_t)r->clips[i] + 1); }
            if (inRegion(r, clipEnd)) { int64_t i = clipEnd - r->start; r->clips[i] = wrap32((int64_t)r->clips[i] + 1); }
            for (int64_t i = 0; i < len; i++) {
                int64_t lp = clipStart + i;
                if (inRegion(r, lp)) region_add(r, lp, bases[readOffset + i], QUAL(readOffset + i), adjMq, 0);
            }
            break; }
        case 5: case 3: break;          /* H, N :207-210 */
        default: r->unknown_ops++;      /* :211-212 println */
        }
        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) readOffset += len;     /* consumesReadBases :214 */
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) refOffset += len;      /* consumesReferenceBases :215 */
    }
#undef QUAL
#undef TRUSTED
    r->readCount += 1;                                                                /* :218 */
    return physCovIncr(r, aStart, aEnd, insert, paired, valid);                        /* :219 */
}

/* BamFile.process loop (BamFile.scala:126-139) inside GenomeRegion.processBam (GenomeRegion.scala:287-300) */
int po_region_add_batch(po_region* r, const pb_batch* b, int frag, int long_read, int32_t* insert_sizes_out) {
    if (long_read < 0 || long_read > 2) return PB_ERR_INVALID;
    const int32_t readsBefore = r->readCount;                                           /* BamFile.scala:120 */
    const int64_t baseCountBefore = r->baseCount;                                       /* :121 */
    const int64_t covBeforeBam = roundDivL(r->baseCount, r->size);                      /* :122; PileUpRegion.scala:36 */
    if (frag) for (int64_t i = 0; i < r->size; i++) r->covBefore[i] = wrap32(depthAt(r, i));   /* :290-293 */
    int32_t maxlen = 0;
    for (int64_t rd = 0; rd < b->n_reads; rd++) if (b->read_len[rd] > maxlen) maxlen = b->read_len[rd];
    uint8_t* bases = (uint8_t*)malloc((size_t)maxlen + 1);
    uint8_t* quals = (uint8_t*)malloc((size_t)maxlen + 1);
    int64_t cursor = 0;
    for (int64_t rd = 0; rd < b->n_reads; rd++) {
        unpack_read(b, rd, bases, quals, &cursor);
        int32_t ins = addRead(r, b, rd, bases, quals, long_read);
        if (insert_sizes_out) insert_sizes_out[rd] = ins;
    }
    free(bases); free(quals);
    if (frag) for (int64_t i = 0; i < r->size; i++)
        r->fragCov[i] = wrap32((int64_t)r->fragCov[i] + wrap32(depthAt(r, i)) - r->covBefore[i]);   /* :296-298 */
    if (r->n_bams < 256) {
        r->bam_reads[r->n_bams] = r->readCount - readsBefore;                           /* :143 */
        r->bam_bases[r->n_bams] = r->baseCount - baseCountBefore;                       /* :146 */
        r->bam_cov[r->n_bams] = roundDivL(r->baseCount, r->size) - covBeforeBam;        /* :142 */
    }
    r->n_bams++;
    return PB_OK;
}

/* ---- indel evidence grouping ------------------------------------------------------------ */
static int ev_cmp(const void* a, const void* b) {
    const po_event* x = (const po_event*)a; const po_event* y = (const po_event*)b;
    if (x->loc != y->loc) return x->loc < y->loc ? -1 : 1;
    if (x->kind != y->kind) return x->kind < y->kind ? -1 : 1;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    return memcmp(x->bytes, y->bytes, (size_t)x->len);
}

typedef struct { int32_t loc, kind, list_len, win_count, win_len, win_has_n; const uint8_t* win; } po_group;

static po_group* group_events(po_region* r, int64_t* n_out) {
    qsort(r->ev, (size_t)r->n_ev, sizeof(po_event), ev_cmp);
    po_group* g = (po_group*)xcalloc((size_t)r->n_ev + 1, sizeof(po_group));
    int64_t ng = 0, i = 0;
    while (i < r->n_ev) {
        int64_t j = i;
        po_group* cur = &g[ng++];
        cur->loc = r->ev[i].loc; cur->kind = r->ev[i].kind; cur->win_count = 0;
        while (j < r->n_ev && r->ev[j].loc == cur->loc && r->ev[j].kind == cur->kind) {
            int64_t k = j;
            while (k < r->n_ev && ev_cmp(&r->ev[k], &r->ev[j]) == 0) k++;
            if ((int32_t)(k - j) > cur->win_count) {   /* any maximal entry; strict majority makes ties moot, PileUp.scala:219-220 */
                cur->win_count = (int32_t)(k - j); cur->win = r->ev[j].bytes; cur->win_len = r->ev[j].len;
            }
            j = k;
        }
        cur->list_len = (int32_t)(j - i);
        cur->win_has_n = memchr(cur->win, 'N', (size_t)cur->win_len) != NULL;            /* PileUp.scala:222 */
        i = j;
    }
    *n_out = ng;
    return g;
}

static const po_group* find_group(const po_group* g, int64_t ng, int32_t loc, int32_t kind) {
    int64_t lo = 0, hi = ng;
    while (lo < hi) {
        int64_t m = (lo + hi) / 2;
        if (g[m].loc < loc || (g[m].loc == loc && g[m].kind < kind)) lo = m + 1; else hi = m;
    }
    return (lo < ng && g[lo].loc == loc && g[lo].kind == kind) ? &g[lo] : NULL;
}

/* ---- PileUp.BaseCall, PileUp.scala:132-247 ------------------------------------------------ */
typedef struct {
    int baseIndex, altBaseIndex, base /*0..3, 4 = N*/, homo, indel /*0,1,2*/, homoIndel, called, highConf;
    int64_t n, score, q, baseSum, altBaseSum;
    int32_t indel_len;
} po_call;

static void order4(const int64_t* s, int* o) {   /* BaseSum.scala:57-60: stable descending */
    o[0] = 0; o[1] = 1; o[2] = 2; o[3] = 3;
    for (int i = 1; i < 4; i++) { int v = o[i], j = i; while (j > 0 && s[o[j - 1]] < s[v]) { o[j] = o[j - 1]; j--; } o[j] = v; }
}

/* hetIndelCall, PileUp.scala:209-247: returns 0 none, 1 homozygous, 2 heterozygous */
static int hetIndelCall(const po_region* r, int64_t depth, const po_group* g, int32_t pct) {
    if (depth < r->cfg.min_min_depth || pct < 5 || g == NULL || g->list_len == 0) return 0;      /* :213 */
    if (g->win_count < 2 || g->win_count <= g->list_len / 2) return 0;                           /* :220 */
    if (g->win_has_n) return 0;                                                                  /* :222 */
    int32_t wl = g->win_len;
    if (r->cfg.old_indel) return (pct >= 33 && pct >= 50 - wl) ? 1 : 0;                          /* :223-228 */
    int32_t middle = 45 - wl > 10 ? 45 - wl : 10;                                                /* :232 */
    int32_t low = middle / 2, high = middle + middle - low;                                      /* :234-236 */
    if (pct > high) return 1;
    if (pct >= low) return 2;
    return 0;
}

static void baseCall(const po_region* r, int64_t i, const po_group* g, int64_t ng, po_call* c) {
    const int64_t* bcnt = r->bc + 4 * i; const int64_t* qs = r->qs + 4 * i;
    int64_t n = bcnt[0] + bcnt[1] + bcnt[2] + bcnt[3];                                           /* :133 */
    int32_t mqSum = r->mqSum[i], qSum = r->qSum[i];
    int32_t ins = r->insertions[i], del = r->deletions[i];
    int64_t depth = n + del;
    int o[4];
    order4(qSum > 0 ? qs : bcnt, o);                                                             /* :135 */
    c->n = n; c->baseIndex = o[0]; c->altBaseIndex = o[1];
    c->base = n > 0 ? o[0] : 4;                                                                  /* :138 */
    c->baseSum = qs[o[0]]; c->altBaseSum = qs[o[1]];
    int64_t total = qs[0] + qs[1] + qs[2] + qs[3];
    int64_t homoScore = c->baseSum - (total - c->baseSum);                                       /* :144 */
    int64_t half = total / 2;
    int64_t heteroScore = total - llabs(half - c->baseSum) - llabs(half - c->altBaseSum);        /* :146 */
    c->homo = homoScore >= heteroScore;
    c->score = mqSum > 0 ? (int64_t)((uint64_t)llabs(homoScore - heteroScore) * (uint64_t)n) / mqSum : 0;  /* :148 */
    c->indel = 0; c->homoIndel = 1; c->indel_len = 0;
    int res = 0;
    if (ins > 2 && ins > del) {                                                                  /* :183-186 */
        int32_t p1 = pctI(r->insQual[i], mqSum), p2 = pctI(ins, wrap32(n));                      /* :122 */
        const po_group* gg = find_group(g, ng, (int32_t)i, 1);
        res = hetIndelCall(r, depth, gg, p1 > p2 ? p1 : p2);
        if (res) { c->indel = 1; c->homoIndel = res == 1; c->indel_len = gg->win_len; }
    }
    if (!res && del > 2 && del > ins) {                                                          /* :188-191 */
        int32_t p1 = pctI(r->delQual[i], mqSum), p2 = pctI(del, wrap32((int64_t)wrap32(n) + del));   /* :123 */
        const po_group* gg = find_group(g, ng, (int32_t)i, 2);
        res = hetIndelCall(r, depth, gg, p1 > p2 ? p1 : p2);
        if (res) { c->indel = 2; c->homoIndel = res == 1; c->indel_len = gg->win_len; }
    }
    c->called = (c->base != 4) || c->indel;                                                      /* :165 */
    c->q = n > 0 ? c->score / n : 0;                                                             /* :166 */
    c->highConf = c->q >= 10;                                                                    /* :167 */
}

static uint64_t pack_call(const po_call* c) {
    return (uint64_t)c->base | ((uint64_t)c->altBaseIndex << 3) | ((uint64_t)c->homo << 5) |
           ((uint64_t)c->indel << 6) | ((uint64_t)c->homoIndel << 8) | ((uint64_t)c->called << 9) |
           ((uint64_t)c->highConf << 10) | ((uint64_t)c->score << 16);
}

/* PileUpRegion.postProcess + GenomeRegion.postProcess pass 1 */
int po_region_finish(po_region* r, pb_region_result* res) {
    int64_t S = r->size;
    /* computePhysCov, PileUpRegion.scala:90-100 */
    r->physCov[0] = wrap32((int64_t)r->physCov[0] + r->physCovStart);
    r->insertSize[0] = wrap32((int64_t)r->insertSize[0] + r->insertSizeStart);
    for (int64_t i = 1; i < S; i++) {
        r->physCov[i] = wrap32((int64_t)r->physCov[i] + r->physCov[i - 1]);
        r->insertSize[i] = wrap32((int64_t)r->insertSize[i] + r->insertSize[i - 1]);
    }
    for (int64_t i = 0; i < S; i++) if (r->physCov[i] > 0) r->insertSize[i] /= r->physCov[i];

    int64_t meanCoverage = roundDivL(r->baseCount, S);                                  /* PileUpRegion.scala:36 */
    int32_t minDepth;
    if (r->cfg.min_depth >= 1) minDepth = (int32_t)r->cfg.min_depth;                     /* GenomeRegion.scala:221-224 */
    else {
        double v = floor(r->cfg.min_depth * (double)meanCoverage + 0.5);                 /* Double.round */
        minDepth = (int32_t)v > r->cfg.min_min_depth ? (int32_t)v : r->cfg.min_min_depth;
    }
    int64_t ng = 0;
    po_group* g = group_events(r, &ng);
    uint8_t* flags = (uint8_t*)xcalloc((size_t)S, 1);
    int32_t* cov = (int32_t*)xcalloc((size_t)S, 4);
    int8_t* wq = (int8_t*)xcalloc((size_t)S, 1);
    int8_t* wmq = (int8_t*)xcalloc((size_t)S, 1);
    int fixamb = r->cfg.fix_amb;
    if (r->readCount != 0) {                                                             /* :229-231 */
        for (int64_t i = 0; i < S; i++) {                                                /* :237-272 */
            po_call c; baseCall(r, i, g, ng, &c);
            int64_t n = depthAt(r, i);
            int rb = r->contig[(int64_t)r->start + i - 1];
            if (rb >= 'a' && rb <= 'z') rb -= 32;                                        /* :783-787 toUpper */
            int rbi = rb == 'A' ? 0 : rb == 'C' ? 1 : rb == 'G' ? 2 : rb == 'T' ? 3 : (rb == 'N' ? 4 : 5);
            cov[i] = wrap32(n);
            int64_t qsum = r->qs[4 * i] + r->qs[4 * i + 1] + r->qs[4 * i + 2] + r->qs[4 * i + 3];
            wq[i] = (int8_t)(uint8_t)roundDivL(qsum, r->mqSum[i]);                       /* :251, PileUp.scala:60-62 */
            wmq[i] = (int8_t)(uint8_t)roundDivL(qsum, r->qSum[i]);                       /* :252, PileUp.scala:56-58 */
            if (n >= minDepth && rbi != 4 && !(flags[i] & PB_FL_DELETED) && c.called) {  /* :255 */
                int b_eq_r = (c.base == rbi);   /* base 'N' (4) can equal only r == 'N', excluded above */
                if (c.homo && b_eq_r && c.highConf && !c.indel) flags[i] |= PB_FL_CONFIRMED;
                else if (c.indel == 1 && c.homoIndel) flags[i] |= PB_FL_CHANGED | (PB_KIND_INS << PB_FL_KIND_SHIFT);
                else if (c.indel == 2 && c.homoIndel) {
                    flags[i] |= PB_FL_CHANGED | (PB_KIND_DEL << PB_FL_KIND_SHIFT);
                    for (int32_t j = 1; j < c.indel_len; j++) {                          /* :260-264 */
                        flags[i + j] |= PB_FL_DELETED;
                        r->deletions[i + j] = wrap32((int64_t)r->deletions[i + j] + r->deletions[i]);
                    }
                } else if (!b_eq_r && c.score > 0) {
                    if (c.homo) flags[i] |= PB_FL_CHANGED | (PB_KIND_SNP << PB_FL_KIND_SHIFT);
                    else if (fixamb || c.altBaseIndex != rbi) flags[i] |= PB_FL_AMBIGUOUS | (PB_KIND_AMB << PB_FL_KIND_SHIFT);
                }
            }
        }
    }
    /* ---- export -------------------------------------------------------------------------- */
    res->size = S; res->base_count = r->baseCount; res->coverage = meanCoverage;
    res->aligned_bases = r->aligned_bases; res->read_count = r->readCount; res->min_depth = minDepth;
    res->unknown_ops = r->unknown_ops; res->dropped_oob = r->dropped_oob;
    if (res->base_count4) for (int64_t i = 0; i < 4 * S; i++) res->base_count4[i] = (int32_t)r->bc[i];
    if (res->qual_sum4) memcpy(res->qual_sum4, r->qs, (size_t)S * 32);
#define CP(dst, src) if (res->dst) memcpy(res->dst, src, (size_t)S * sizeof(*(res->dst)))
    CP(mq_sum, r->mqSum); CP(q_sum, r->qSum); CP(phys_cov, r->physCov); CP(insert_size, r->insertSize);
    CP(bad_pair, r->badPair); CP(deletions, r->deletions); CP(del_qual, r->delQual);
    CP(insertions, r->insertions); CP(ins_qual, r->insQual); CP(clips, r->clips);
    CP(coverage_arr, cov); CP(frag_coverage, r->fragCov); CP(weighted_qual, wq); CP(weighted_mq, wmq);
    CP(flags, flags);
#undef CP
    if (res->call) for (int64_t i = 0; i < S; i++) {   /* final-state BaseCall: what Vcf.writeRecord recomputes, Vcf.scala:78-79 */
        po_call c; baseCall(r, i, g, ng, &c); res->call[i] = pack_call(&c);
    }
    /* the sparse form of the call plane: the loci identifyAndFixIssues looks at (GenomeRegion.scala:307-380) */
    res->n_calls = 0;
    if (res->calls && res->calls_cap > 0) for (int64_t i = 0; i < S; i++) {
        if (!(flags[i] & (PB_FL_CHANGED | PB_FL_AMBIGUOUS))) continue;
        if (res->n_calls < res->calls_cap) {
            po_call c; baseCall(r, i, g, ng, &c);
            pb_call_entry* en = &res->calls[res->n_calls];
            en->locus_index = (int32_t)i; en->flags = flags[i]; en->call = pack_call(&c);
        }
        res->n_calls++;
    }
    int64_t nb = 0, ni = 0;
    for (int64_t k = 0; k < ng; k++) {
        if (res->indels && ni < res->indels_cap) {
            pb_indel* o = &res->indels[ni];
            o->locus_index = g[k].loc; o->kind = g[k].kind; o->list_len = g[k].list_len;
            o->win_count = g[k].win_count; o->win_len = g[k].win_len; o->win_has_n = g[k].win_has_n;
            o->str_off = nb;
            if (res->indel_bytes && nb + g[k].win_len <= res->indel_bytes_cap)
                memcpy(res->indel_bytes + nb, g[k].win, (size_t)g[k].win_len);
        }
        ni++; nb += g[k].win_len;
    }
    res->n_indels = ni; res->n_indel_bytes = nb;
    res->n_batches = r->n_bams;
    for (int b = 0; b < r->n_bams && b < 256 && b < res->batch_cap; b++) {
        if (res->batch_read_count) res->batch_read_count[b] = r->bam_reads[b];
        if (res->batch_base_count) res->batch_base_count[b] = r->bam_bases[b];
        if (res->batch_coverage) res->batch_coverage[b] = r->bam_cov[b];
    }
    free(g); free(flags); free(cov); free(wq); free(wmq);
    return PB_OK;
}
