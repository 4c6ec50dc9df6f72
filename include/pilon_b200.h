/*
 * pilon_b200.h -- C ABI of the B200-native pileup + BaseCall engine (libpilonb200.so).
 *
 * This is the drop-in boundary for ONE hot path of broadinstitute/pilon: everything the Scala
 * class `PileUpRegion` and its per-locus `PileUp` / `PileUp.BaseCall` view compute
 * (reference: src/main/scala/org/broadinstitute/pilon/PileUpRegion.scala, PileUp.scala,
 * BaseSum.scala, Utils.scala) plus pass 1 of `GenomeRegion.postProcess`
 * (GenomeRegion.scala:214-272) and the fragCoverage bookkeeping of `GenomeRegion.processBam`
 * (GenomeRegion.scala:287-300).  The reference has no FFI of its own (pure JVM); the entry
 * points below are what a re-plumbed `BamFile.process` (BamFile.scala:108-148) and a
 * `PileUpRegion` facade would bind through Panama / JNA -- see INTEGRATION.md.
 *
 * Conventions: plain C, no exceptions, no torch types.  Every function returns PB_OK (0) or a
 * negative PB_ERR_* code; pb_last_error() returns a thread-local message for the last failure.
 * An engine handle is NOT thread-safe; use one handle per (GPU, stream).  All citations
 * "File.scala:a-b" refer to the reference tree above.
 */
#ifndef PILON_B200_H
#define PILON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 6

/* ---- status codes ---------------------------------------------------------------------- */
#define PB_OK                 0
#define PB_ERR_INVALID       -1   /* bad argument / call order                                */
#define PB_ERR_CUDA          -2   /* CUDA runtime failure (message has the cudaError string)  */
#define PB_ERR_UNSORTED      -3   /* reads of a batch are not sorted by pos                   */
#define PB_ERR_UNSUPPORTED   -4   /* feature gated off (long-read branches)    */
#define PB_ERR_OOM           -5
#define PB_ERR_HASH          -6   /* 32-bit insertion-hash collision detected (never seen)    */

/* ---- configuration: the `object Pilon` vars the path reads (Pilon.scala:28-73) ---------- */
typedef struct pb_config {
    int32_t min_qual;        /* Pilon.minQual      used PileUp.scala:77            default 0   */
    int32_t min_mq;          /* Pilon.minMq        used PileUpRegion.scala:107     default 0   */
    int32_t flank;           /* Pilon.flank        used PileUpRegion.scala:34,118  default 10  */
    int32_t default_qual;    /* Pilon.defaultQual  used PileUpRegion.scala:115     default 10  */
    int32_t min_min_depth;   /* Pilon.minMinDepth  PileUp.scala:213, GenomeRegion.scala:223  5 */
    int32_t old_indel;       /* Pilon.oldIndel     used PileUp.scala:223           default 0   */
    int32_t fix_amb;         /* Pilon.iupac || Pilon.fixAmb  GenomeRegion.scala:235 default 0  */
    int32_t reserved;
    double  min_depth;       /* Pilon.minDepth     used GenomeRegion.scala:221-224 default 0.1 */
} pb_config;

/* ---- one batch of reads, struct-of-arrays ------------------------------------------------
 * What `BamFile.process` (BamFile.scala:126-139) would have fed to `PileUpRegion.addRead` one
 * SAMRecord at a time, for reads that passed `validateRead` (BamFile.scala:101-105).
 * Reads MUST be sorted by `pos` ascending (a coordinate-sorted BAM query is).
 *
 * Per-read flag bits (from SAMRecord accessors used at PileUpRegion.scala:103-111): */
#define PB_F_PAIRED         0x01  /* getReadPairedFlag                                        */
#define PB_F_PROPER         0x02  /* getProperPairFlag                                        */
#define PB_F_MATE_SAME_REF  0x04  /* getReferenceIndex == getMateReferenceIndex               */
#define PB_F_HAS_QUALS      0x08  /* getBaseQualities.size > 0 (BAM: first qual byte != 0xFF) */
#define PB_F_UNMAPPED       0x10  /* getReadUnmappedFlag (getAlignmentEnd == 0)               */
#define PB_F_REVERSE        0x20  /* getReadNegativeStrandFlag (BamFile.scala:137 only)       */

/* CIGAR ops are BAM-encoded: len << 4 | op, op = 0..8 for "MIDNSHP=X". */

/* Base / quality packing.  Each read owns a run of `read_len` bases starting at base index
 * seq_off[r] (a multiple of 4) inside the batch:
 *   bases2[i >> 2] >> (2 * (i & 3)) & 3   = 0,1,2,3 for A,C,G,T
 *   quals[i]                               = phred quality, 0..127, or exactly 0x80 for a base the
 *                                            per-locus counters can never count: a read byte that
 *                                            is not exactly 'A','C','G','T' (PileUp.scala:46-52) or
 *                                            a quality byte >= 128 (a negative JVM Byte).
 * Every 0x80 base has one entry in the sparse exception table (sorted by base index) holding its
 * original ASCII letter and raw quality byte, so the rare paths that need them (indel strings,
 * indel anchor qualities, left-shift comparisons) stay bit-exact.
 * `quals` and `bases2` must be 16-byte aligned and readable for 16 bytes past their last element
 * (the kernels fetch them in aligned 16-byte blocks); the engine's own H2D staging guarantees it.
 * For reads without PB_F_HAS_QUALS the quality bytes are 0x00 / 0x80 only. */
#define PB_MEM_HOST    0   /* pointers are host memory (pinned preferred); engine copies H2D   */
#define PB_MEM_DEVICE  1   /* pointers are device memory on the engine's GPU; used in place    */

typedef struct pb_batch {
    int64_t n_reads;
    int64_t n_cigar;            /* total CIGAR ops                                            */
    int64_t n_seq;              /* total base slots incl. padding; multiple of 4; < 2^32      */
    int64_t n_exc;              /* exception table entries                                    */
    const int32_t*  pos;        /* [n_reads]   getAlignmentStart (1-based)                    */
    const int32_t*  tlen;       /* [n_reads]   getInferredInsertSize                          */
    const int32_t*  read_len;   /* [n_reads]   getReadLength (SEQ length)                     */
    const uint8_t*  mapq;       /* [n_reads]   getMappingQuality                              */
    const uint8_t*  flags;      /* [n_reads]   PB_F_*                                         */
    const uint32_t* cigar_off;  /* [n_reads+1] CSR offsets into cigar                         */
    const uint32_t* cigar;      /* [n_cigar]                                                  */
    const uint32_t* seq_off;    /* [n_reads]   first base index of the read (multiple of 4)   */
    const uint8_t*  quals;      /* [n_seq]                                                    */
    const uint8_t*  bases2;     /* [n_seq / 4]                                                */
    const uint32_t* exc_idx;    /* [n_exc] sorted base indices                                */
    const uint8_t*  exc_base;   /* [n_exc] original ASCII read byte                           */
    const uint8_t*  exc_qual;   /* [n_exc] original quality byte                              */
    int32_t mem;                /* PB_MEM_HOST | PB_MEM_DEVICE                                */
    int32_t qual_code_bits;     /* 0: no packed qualities; 3 or 4: width of the codes in qual_codes */
    /* Optional compact transport of `quals` for PB_MEM_HOST batches (the end-to-end path is PCIe-bound):
     * when qual_codes is non-NULL the engine uploads these ceil(n_seq * qual_code_bits / 8) bytes instead
     * of `quals` and expands them on the device: code i occupies bits [i * bits, (i + 1) * bits) of the
     * little-endian bit stream and quals[i] = qual_lut[code i].  Possible whenever a batch uses at most
     * 8 (3-bit) or 16 (4-bit) distinct quality bytes -- binned instrument qualities; pb_packer_view fills
     * it in automatically with the narrowest width, else leaves it NULL.  `quals` may then be NULL for
     * the engine.  The buffer must be readable up to the next multiple of 16 bytes. */
    const uint8_t*  qual_codes; /* packed codes or NULL                                       */
    uint8_t qual_lut[16];       /* code -> quality byte (0..127, or 0x80)                     */
    /* Optional compact transport of `bases2` for PB_MEM_HOST batches: only the bases that differ from what
     * the reference predicts (the CRAM idea).  The predicted code of a read base is the code of the
     * reference byte an M/=/X operation aligns it to -- when that byte is exactly 'A','C','G','T' and its
     * locus lies in [region start - PB_REF_HALO, region stop + PB_REF_HALO] and in the contig -- and 0
     * otherwise (inserted and soft-clipped bases, any other reference byte, padding).  base_delta_idx
     * lists, sorted, the batch base indices whose stored code differs from the prediction and
     * base_delta_code their stored codes (one per byte).  When base_delta_idx is non-NULL the engine
     * uploads these 5 bytes per entry instead of `bases2` and rebuilds the array on the device;
     * `bases2` may then be NULL for the engine.  pb_base_delta_encode computes them. */
    const uint32_t* base_delta_idx;   /* [n_base_delta] or NULL                               */
    const uint8_t*  base_delta_code;  /* [n_base_delta]                                       */
    int64_t n_base_delta;
    /* Optional compact transport of the per-read arrays (pos, tlen, read_len, mapq, flags, cigar_off, cigar,
     * seq_off: 22 bytes per read + 4 per CIGAR op) for PB_MEM_HOST batches of short reads: one 8-byte record per
     * read, little endian,
     *   bits  0..15  pos - pos of the previous read (of meta_pos0 for the first read); 0xFFFF: see meta_esc
     *   bits 16..31  tlen as int16; -32768: see meta_esc
     *   bits 32..39  read_len (0..255)      bits 40..47  mapq      bits 48..55  flags
     *   bits 56..63  0: the CIGAR is one M over the whole read; n > 0: its n ops are the next n of meta_cigar
     * meta_esc lists, sorted by read index, the values that do not fit: triples (read index, field, value) with
     * field 0 = pos delta, 1 = tlen.  seq_off is implied: read r starts at r * meta_seq_stride when that is > 0,
     * else at the sum of the earlier reads' lengths rounded up to multiples of 4.  When meta_codes is non-NULL
     * the engine uploads these arrays instead of the eight plain ones (which may then be NULL for the engine) and
     * rebuilds the plain ones on the device.  pb_meta_encode computes them, or says that the batch cannot be
     * put this way (reads longer than 255 bases, more than 255 CIGAR ops, another seq_off layout, unsorted). */
    const uint64_t* meta_codes;       /* [n_reads] or NULL                                    */
    const uint32_t* meta_cigar;       /* [n_meta_cigar]                                       */
    const int32_t*  meta_esc;         /* [3 * n_meta_esc]                                     */
    int64_t n_meta_cigar;
    int64_t n_meta_esc;
    int32_t meta_pos0;
    int32_t meta_seq_stride;
} pb_batch;

#define PB_REF_HALO 16384   /* loci of reference kept on the device on either side of a region */

/* ---- per-locus call record (PileUp.BaseCall, PileUp.scala:132-167) -----------------------
 * Computed on the FINAL per-locus state, i.e. after pass 1 spilled homozygous-deletion counts
 * into the deleted loci (GenomeRegion.scala:259-264) -- what Vcf.writeRecord would see.
 *   bits  0..2  base      0..3 = A,C,G,T ; 4 = 'N' (n == 0)          (:138)
 *   bits  3..4  altBase   0..3                                        (:140)
 *   bit   5     homo                                                  (:147)
 *   bits  6..7  indel     0 none, 1 insertion, 2 deletion             (:151-162)
 *   bit   8     homoIndel                                             (:151-162)
 *   bit   9     called                                                (:165)
 *   bit   10    highConfidence (q >= 10)                              (:167)
 *   bits 16..63 score (48 bit, non-negative)                          (:148)            */
#define PB_CALL_BASE(c)      ((int)((c) & 7))
#define PB_CALL_ALT(c)       ((int)(((c) >> 3) & 3))
#define PB_CALL_HOMO(c)      ((int)(((c) >> 5) & 1))
#define PB_CALL_INDEL(c)     ((int)(((c) >> 6) & 3))
#define PB_CALL_HOMOINDEL(c) ((int)(((c) >> 8) & 1))
#define PB_CALL_CALLED(c)    ((int)(((c) >> 9) & 1))
#define PB_CALL_HICONF(c)    ((int)(((c) >> 10) & 1))
#define PB_CALL_SCORE(c)     ((int64_t)((c) >> 16))

/* Pass-1 disposition flags (GenomeRegion.scala:45-49, 84-88, 255-271) */
#define PB_FL_CONFIRMED  0x01
#define PB_FL_CHANGED    0x02   /* SNP, INS or DEL                                            */
#define PB_FL_AMBIGUOUS  0x04   /* AMB                                                        */
#define PB_FL_DELETED    0x08
#define PB_FL_KIND_SHIFT 4      /* bits 4..5: 0 SNP, 1 INS, 2 DEL, 3 AMB (valid if CHANGED|AMBIGUOUS) */
#define PB_KIND_SNP 0
#define PB_KIND_INS 1
#define PB_KIND_DEL 2
#define PB_KIND_AMB 3

/* One entry per (locus, kind) that received at least one insertion / deletion. */
typedef struct pb_indel {
    int32_t  locus_index;   /* 0-based index into the region                                  */
    int32_t  kind;          /* 1 insertion, 2 deletion                                        */
    int32_t  list_len;      /* insertionList / deletionList length (PileUp.scala:40-41)       */
    int32_t  win_count;     /* occurrences of the most frequent string                        */
    int32_t  win_len;       /* its length                                                     */
    int32_t  win_has_n;     /* contains 'N' (PileUp.scala:222)                                */
    int64_t  str_off;       /* offset of its bytes in pb_region_result.indel_bytes            */
} pb_indel;

/* ---- region results ------------------------------------------------------------------------
 * All array pointers are HOST memory owned by the caller, each `size` elements long unless noted;
 * a NULL pointer means "not wanted" (no device->host copy is made for it). */
typedef struct pb_call_entry {
    int32_t  locus_index;    /* 0-based in the region                                          */
    uint32_t flags;          /* PB_FL_* of the locus                                           */
    uint64_t call;           /* packed BaseCall record, as in the `call` plane                 */
} pb_call_entry;

typedef struct pb_region_result {
    /* scalars, always filled */
    int64_t size;            /* stop + 1 - start                               Region.scala:27 */
    int64_t base_count;      /* PileUpRegion.baseCount                   PileUpRegion.scala:32 */
    int64_t coverage;        /* roundDiv(baseCount, size)                PileUpRegion.scala:36 */
    int64_t aligned_bases;   /* sum of M/=/X lengths of the submitted reads (bench metric)     */
    int32_t read_count;      /* PileUpRegion.readCount                   PileUpRegion.scala:33 */
    int32_t min_depth;       /* GenomeRegion.minDepth                GenomeRegion.scala:221-224 */
    int32_t unknown_ops;     /* CIGAR ops that hit the println at    PileUpRegion.scala:211-212 */
    int32_t dropped_oob;     /* indels whose left shift left the region (JVM: AIOOBE crash)    */
    int64_t n_indels;        /* entries written to `indels`                                    */
    int64_t n_indel_bytes;   /* bytes written to `indel_bytes`                                 */

    /* PileUp counters (PileUp.scala:26-39), final state */
    int32_t* base_count4;    /* [size*4] baseCount.sums, locus-major (values fit 31 bits)      */
    int64_t* qual_sum4;      /* [size*4] qualSum.sums, locus-major                             */
    int32_t* mq_sum;
    int32_t* q_sum;
    int32_t* phys_cov;       /* after computePhysCov                 PileUpRegion.scala:90-100 */
    int32_t* insert_size;    /* after computePhysCov                                           */
    int32_t* bad_pair;
    int32_t* deletions;      /* including the pass-1 spill          GenomeRegion.scala:263     */
    int32_t* del_qual;
    int32_t* insertions;
    int32_t* ins_qual;
    int32_t* clips;

    /* GenomeRegion pass-1 arrays that are not plain copies of the above */
    int32_t* coverage_arr;   /* coverage(i) = depth.toInt           GenomeRegion.scala:247     */
    int32_t* frag_coverage;  /* fragCoverage(i)                     GenomeRegion.scala:296-298 */
    int8_t*  weighted_qual;  /* weightedQual.toByte                 GenomeRegion.scala:251     */
    int8_t*  weighted_mq;    /* weightedMq.toByte                   GenomeRegion.scala:252     */
    uint8_t* flags;          /* PB_FL_*                                                        */
    uint64_t* call;          /* packed BaseCall record                                         */

    /* sparse indel evidence; capacities are inputs, counts come back in n_indels/n_indel_bytes */
    pb_indel* indels;        int64_t indels_cap;
    uint8_t*  indel_bytes;   int64_t indel_bytes_cap;

    /* per-BAM deltas, one entry per pb_region_add_batch call in call order: what BamFile.process brackets its read
     * loop with (BamFile.scala:120-122,142-146).  batch_cap is an input (entries the three arrays can hold; arrays
     * may be NULL), n_batches comes back.
     *   batch_read_count[b]  readCount after - before            (:121,143)
     *   batch_base_count[b]  baseCount after - before, what BamFile.baseCount accumulates for
     *                        GenomeFile.coverageSummary            (:120,146; GenomeFile.scala:178-187)
     *   batch_coverage[b]    coverage after - coverage before, the value process returns (:122,142,147)  */
    int32_t* batch_read_count;
    int64_t* batch_base_count;
    int64_t* batch_coverage;
    int64_t  batch_cap;
    int64_t  n_batches;

    /* The call plane, sparse: one entry per locus whose pass-1 call changes or questions the reference
     * (flags & (PB_FL_CHANGED | PB_FL_AMBIGUOUS)), ascending locus_index -- the only loci whose call record
     * identifyAndFixIssues reads (GenomeRegion.scala:307-380).  A `--fix snps,indels --changes` caller asks for
     * `flags`, `frag_coverage` (pass 2, :275-283) and these entries instead of the 8-byte-per-locus `call` plane.
     * calls_cap is an input (NULL / 0 = not wanted); n_calls comes back and counts every such locus, also the
     * ones beyond the capacity (which are not written).  */
    pb_call_entry* calls;
    int64_t  calls_cap;
    int64_t  n_calls;
} pb_region_result;

typedef struct pb_engine pb_engine;

/* library-level */
int         pb_abi_version(void);
const char* pb_last_error(void);
int         pb_device_count(int* n_out);

/* Engine lifetime = what GenomeRegion.initializePileUps / finalizePileUps bracket
 * (GenomeRegion.scala:149-155), but reusable across regions to keep device buffers warm. */
int pb_create(int device, const pb_config* cfg, pb_engine** out);
int pb_destroy(pb_engine* e);

/* new PileUpRegion(name, start, stop)  (GenomeRegion.scala:150; PileUpRegion.scala:26-36).
 * contig_bases = GenomeRegion.contigBases (raw FASTA bytes, case preserved), host memory. */
int pb_region_begin(pb_engine* e, const uint8_t* contig_bases, int64_t contig_len,
                    int32_t start, int32_t stop);

/* The batched equivalent of the `for (read <- reads) ... addRead` loop (BamFile.scala:126-139).
 * counts_toward_frag_coverage = (bamType != "jumps")      (GenomeRegion.scala:291,296)
 * long_read_type              = BamFile.longReadType       (BamFile.scala:43-47): 0, 1 = nanopore, 2 = pacbio; the
 *                               long-read branches of addRead (PileUpRegion.scala:120-134,142,160,180-181,190) apply
 *                               to the reads of this batch.
 * The call is asynchronous on the engine's stream; host buffers must stay valid until
 * pb_region_finish returns. */
int pb_region_add_batch(pb_engine* e, const pb_batch* batch,
                        int counts_toward_frag_coverage, int long_read_type);

/* PileUpRegion.postProcess + GenomeRegion.postProcess pass 1 (PileUpRegion.scala:226-229;
 * GenomeRegion.scala:214-272) and the copy of everything the driver / writers read back.
 * insert_sizes_out[b] (may be NULL) receives, for batch b in add order, the per-read return value
 * of addRead (PileUpRegion.scala:219) as int32[n_reads] -- what BamFile.addInsert consumes. */
int pb_region_finish(pb_engine* e, pb_region_result* res, int32_t* const* insert_sizes_out);

/* Device-resident timing hook for bench.py: (re)runs the compute part of pb_region_finish on the
 * batches already added, without any host<->device copy, `iters` times, and reports the elapsed
 * device milliseconds measured with CUDA events on the engine's stream (total and for the pileup
 * kernel alone), plus the number of kernel launches. */
int pb_region_compute_timed(pb_engine* e, int iters, float* total_ms, float* pileup_ms,
                            int64_t* launches);

/* Same compute pass, launched on the engine's stream without timing and without waiting for it to
 * finish: the pass has no host round trip (sizes and scalars stay on the device), so it is enqueued at
 * once; from the third call on the same batch set it is replayed as one captured CUDA graph
 * (PB_NOGRAPH=1 disables that).  Lets a caller keep several engines -- one per host thread -- busy on
 * one GPU so that the small kernels of one region overlap the large kernels of another.  Error flags
 * of such a pass are only reported by pb_region_finish / pb_region_compute_timed.  Synchronise the
 * stream (pb_stream) before reading anything. */
int pb_region_compute(pb_engine* e);

/* Base deltas of a host batch (see pb_batch.base_delta_idx) against `contig` for the region [start, stop]
 * the batch will be added to.  *idx_out / *code_out are malloc'ed (release with pb_free), *n_out entries. */
int pb_base_delta_encode(const pb_batch* b, const uint8_t* contig, int64_t contig_len, int32_t start, int32_t stop,
                         uint32_t** idx_out, uint8_t** code_out, int64_t* n_out);
void pb_free(void* p);

/* Compact per-read metadata of a host batch (see pb_batch.meta_codes).  Outputs are malloc'ed (release with pb_free).
 * Returns PB_ERR_UNSUPPORTED, leaving the outputs untouched, when the batch cannot be put this way. */
int pb_meta_encode(const pb_batch* b, uint64_t** codes_out, uint32_t** cigar_out, int64_t* n_cigar_out,
                   int32_t** esc_out, int64_t* n_esc_out, int32_t* pos0_out, int32_t* seq_stride_out);

/* Raw CUDA stream handle (cudaStream_t) so callers can order their own copies against the engine. */
int pb_stream(pb_engine* e, void** stream_out);

/* ---- host-side packer (the `BamFile.process` side of the seam) ---------------------------
 * Packs one read given the way BAM stores it / htsjdk exposes it into the arrays of a pb_batch
 * under construction.  `seq` is ASCII (getReadBases), `qual` raw phred bytes or NULL. */
typedef struct pb_packer pb_packer;
int pb_packer_create(pb_packer** out);
int pb_packer_destroy(pb_packer* p);
int pb_packer_reset(pb_packer* p);
int pb_packer_add(pb_packer* p, int32_t pos, int32_t tlen, int32_t mapq, uint32_t flags,
                  const uint32_t* cigar, int32_t n_cigar,
                  const uint8_t* seq, const uint8_t* qual, int32_t read_len);
/* The same for a record in BAM's own encoding: `seq4` = 4-bit bases, two per byte, high nibble first
 * (=ACMGRSVTWYHKDBN), as they lie in a BAM record -- what pb_bam_query_pack feeds the packer with. */
int pb_packer_add_bam(pb_packer* p, int32_t pos, int32_t tlen, int32_t mapq, uint32_t flags,
                      const uint32_t* cigar, int32_t n_cigar,
                      const uint8_t* seq4, const uint8_t* qual, int32_t read_len);
/* Bulk form: reads already in struct-of-arrays with ASCII bases (one byte per base, unpadded). */
int pb_packer_add_many(pb_packer* p, int64_t n_reads, const int32_t* pos, const int32_t* tlen,
                       const uint8_t* mapq, const uint8_t* flags, const int32_t* read_len,
                       const uint32_t* cigar_off, const uint32_t* cigar,
                       const uint64_t* ascii_off, const uint8_t* seq, const uint8_t* qual);
/* The returned view points into the packer's (pinned when possible) buffers, valid until the next
 * reset / add / destroy. */
int pb_packer_view(pb_packer* p, pb_batch* out);

/* ---- consumers of the per-locus results (host side) -----------------------
 * GenomeRegion.postProcess pass 2 (GenomeRegion.scala:275-283), identifyAndFixIssues for `--fix snps,indels`
 * (:307-380,413), fixFixList / fixIssues (:557-621), writeChanges (:646-657), writeVcf (:623-643) with
 * Vcf.writeRecord / writeDup (Vcf.scala:74-176,193-201), the wiggle tracks (Tracks.scala:56-186) and GenomeFile's
 * naming / FASTA rules (GenomeFile.scala:79-82,137-141).  Gap filling and local reassembly stay the reference's
 * CPU code (north_star): bigFixList is empty here. */
typedef struct pb_output_config {      /* the `object Pilon` switches these consumers read (Pilon.scala:28-73) */
    int32_t fix_snps;        /* Pilon.fixSnps    --fix snps                                            */
    int32_t fix_indels;      /* Pilon.fixIndels  --fix indels                                          */
    int32_t iupac;           /* Pilon.iupac      AMB fixes become IUPAC codes     GenomeRegion.scala:333 */
    int32_t diploid;         /* Pilon.diploid    no Amb filter, one snp total     Vcf.scala:122         */
    int32_t vcf_qe;          /* Pilon.vcfQE      QE= instead of QP=               Vcf.scala:145         */
    int32_t longread;        /* Pilon.longread   AMB calls are not fixed          GenomeRegion.scala:332 */
    int32_t reserved[2];
} pb_output_config;

typedef struct pb_out_stats {          /* the numbers of the region log lines (GenomeRegion.scala:358-370) */
    int64_t confirmed, non_n, snps, amb, ins, dels, ins_bases, del_bases;
    int64_t n_fixes;         /* entries of fixFixList(snpFixList ++ smallFixList)                        */
    int64_t fix_mismatches;  /* "Fix mismatch" log lines (GenomeRegion.scala:605-606,614)                */
    int64_t n_dups;          /* duplicationEvents (:735-741)                                             */
} pb_out_stats;

/* Tracks.scala track functions that are plain per-locus maps of the engine's planes */
#define PB_TRACK_CHANGES            0   /* Tracks.scala:57-60   */
#define PB_TRACK_UNCONFIRMED        1   /* :62-65               */
#define PB_TRACK_COPY_NUMBER        2   /* :67-70               */
#define PB_TRACK_COVERAGE           3   /* :77-80               */
#define PB_TRACK_BAD_COVERAGE       4   /* :93-96               */
#define PB_TRACK_PCT_BAD            5   /* :139-147             */
#define PB_TRACK_DELTA_COVERAGE     6   /* :104-107             */
#define PB_TRACK_DIP_COVERAGE       7   /* :109-112             */
#define PB_TRACK_PHYSICAL_COVERAGE  8   /* :114-117             */
#define PB_TRACK_CLIPPED            9   /* :163-166             */
#define PB_TRACK_WEIGHTED_QUAL     10   /* :157-160             */
#define PB_TRACK_WEIGHTED_MQ       11   /* :150-155             */

typedef struct pb_region_out pb_region_out;

const char* pb_out_last_error(void);
/* One GenomeRegion downstream of postProcess pass 1.  `res` (with flags, call, frag_coverage and untruncated indel
 * evidence) and `contig_bases` are referenced, not copied: both must outlive the handle. */
int pb_out_create(const pb_region_result* res, const uint8_t* contig_bases, int64_t contig_len, const char* name,
                  int32_t start, int32_t stop, const pb_output_config* cfg, pb_region_out** out);
int pb_out_destroy(pb_region_out* o);
int pb_out_stats_get(const pb_region_out* o, pb_out_stats* st);
int pb_out_bases(const pb_region_out* o, const uint8_t** bases, int64_t* n);        /* GenomeRegion.bases after fixIssues    */
int pb_out_copy_number(const pb_region_out* o, const int16_t** cn, int64_t* n);    /* GenomeRegion.copyNumber               */
/* Text producers: *text points into a buffer owned by the handle, valid until the same function is called again. */
int pb_out_log(pb_region_out* o, const char** text, int64_t* n);                    /* the region's log lines               */
int pb_out_changes(pb_region_out* o, const char* new_name, int64_t offset, const char** text, int64_t* n);   /* writeChanges */
int pb_out_vcf(pb_region_out* o, int threads, const char** text, int64_t* n);       /* writeVcf: needs every counter plane  */
int pb_out_wig(pb_region_out* o, int track, const char** text, int64_t* n);         /* makeTrack body for this region       */
/* GenomeFile-level helpers; buf == NULL only reports the size in *n. */
int pb_pilon_name(const char* name, char* buf, int64_t cap, int64_t* n);
int pb_fasta_element(const char* header, const uint8_t* bases, int64_t n_bases, char* buf, int64_t cap, int64_t* n);
int pb_vcf_header(const pb_output_config* cfg, const char* date, const char* version, const char* command_args,
                  const char* reference_uri, const char* const* contig_names, const int64_t* contig_sizes, int32_t n_contigs,
                  char* buf, int64_t cap, int64_t* n);

/* ---- BAM ingest and synthetic BAM / BAI / FASTA output (host side) ---------
 * pb_bam_query_pack is BamFile.process's reader loop (BamFile.scala:117-139) without htsjdk: BGZF inflate, BAI
 * linear-index seek, BAM record decode, validateRead (BamFile.scala:101-105) and pb_packer_add.  The writer produces
 * coordinate-sorted BAM + BAI (+ FASTA / .fai) from pb_batch arrays so that synthetic inputs can also be fed to the
 * reference JVM (tools/run_real_pilon.sh). */
typedef struct pb_bam pb_bam;
typedef struct pb_bam_writer pb_bam_writer;
const char* pb_bam_last_error(void);
int pb_bam_open(const char* bam_path, const char* bai_path /* NULL: bam_path + ".bai" */, pb_bam** out);
int pb_bam_close(pb_bam* b);
int pb_bam_n_refs(const pb_bam* b, int32_t* n);
int pb_bam_ref(const pb_bam* b, int32_t i, const char** name, int64_t* len);
/* records of reference ref_id overlapping [start, stop] (1-based inclusive: pass region.start - 10000, region.stop + 10000
 * as BamFile.scala:118-119 does); non_pf = Pilon.nonPf, duplicates = Pilon.duplicates */
int pb_bam_query_pack(pb_bam* b, int32_t ref_id, int32_t start, int32_t stop, int non_pf, int duplicates,
                      pb_packer* packer, int64_t* n_records, int64_t* n_rejected);
int pb_bam_writer_open(const char* path, const char* const* ref_names, const int64_t* ref_lens, int32_t n_refs,
                       const char* program_line, pb_bam_writer** out);
int pb_bam_writer_add_batch(pb_bam_writer* w, int32_t ref_id, const pb_batch* batch, const uint16_t* extra_flags);
int pb_bam_writer_close(pb_bam_writer* w, const char* bai_path);
int pb_fasta_write(const char* path, const char* const* names, const uint8_t* const* seqs, const int64_t* lens, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* PILON_B200_H */
