// pb_output.cpp -- the consumers of the engine's per-locus results, host side of libpilonb200.so:
// pass 2 of GenomeRegion.postProcess, GenomeRegion.identifyAndFixIssues for `--fix snps,indels`, fixFixList / fixIssues,
// writeChanges, writeVcf + Vcf.writeRecord, the wiggle tracks and GenomeFile's FASTA / naming rules.
//
// Everything here is a pure function of a pb_region_result (what pb_region_finish filled in), the contig bytes and a
// handful of `object Pilon` switches; no GPU is involved (per-locus text formatting is spread over host threads).
// Reference citations are relative to /root/reference/src/main/scala/org/broadinstitute/pilon/.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pilon_b200.h"

namespace {

thread_local std::string g_err_out;
int fail_out(int code, const std::string& msg) { g_err_out = msg; return code; }

inline int32_t wrap32(int64_t x) { return (int32_t)(uint32_t)(uint64_t)x; }
inline int64_t roundDivL(int64_t n, int64_t d) { return d > 0 ? (int64_t)((uint64_t)n + (uint64_t)(d / 2)) / d : 0; }   // Utils.scala:23
inline int32_t roundDivI(int32_t n, int32_t d) { return d > 0 ? wrap32((int64_t)n + d / 2) / d : 0; }                   // Utils.scala:24
inline int32_t pctI(int32_t n, int32_t d) { return roundDivI(wrap32(100LL * n), d); }                                   // Utils.scala:26
inline char upper(uint8_t b) { return (char)((b >= 'a' && b <= 'z') ? b - 32 : b); }

struct Fix { int32_t locus; std::string was, patch; };          // GenomeRegion.Fix

void put_i64(std::string& s, int64_t v) {
    char buf[24]; int n = 0;
    uint64_t u = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) buf[n++] = '-';
    while (n) s.push_back(buf[--n]);
}

char iupac_of(char a, char b) {                                  // Bases.scala:62-88
    auto bit = [](char c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 0; };
    static const char tab[16] = {'?', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};
    return tab[(bit(a) | bit(b)) & 15];
}

}  // namespace

struct pb_region_out {
    const pb_region_result* res = nullptr;
    const uint8_t* contig = nullptr; int64_t contig_len = 0;
    std::string name; int32_t start = 0, stop = 0; int64_t size = 0;
    pb_output_config cfg{};
    std::vector<int16_t> copy_number;                            // GenomeRegion.copyNumber (:56, pass 2 :275-283)
    std::vector<Fix> snp_fixes, small_fixes;                     // newest first, like the Scala lists (:303-305)
    std::vector<uint8_t> bases;                                  // GenomeRegion.bases after fixIssues
    pb_out_stats stats{};
    std::string log, changes, vcf, wig;

    char ref_base(int64_t locus) const { return upper(contig[locus - 1]); }                 // :783-787
    const pb_indel* indel_at(int64_t i, int kind) const {                                   // evidence entry of (locus, kind)
        const pb_indel* a = res->indels; int64_t lo = 0, hi = std::min<int64_t>(res->n_indels, res->indels_cap);
        while (lo < hi) {
            const int64_t m = (lo + hi) >> 1;
            if (a[m].locus_index < i || (a[m].locus_index == i && a[m].kind < kind)) lo = m + 1; else hi = m;
        }
        return (lo < std::min<int64_t>(res->n_indels, res->indels_cap) && a[lo].locus_index == i && a[lo].kind == kind) ? &a[lo] : nullptr;
    }
    std::string indel_string(int64_t i, int kind) const {
        const pb_indel* e = indel_at(i, kind);
        if (!e || !res->indel_bytes) return std::string();
        return std::string(reinterpret_cast<const char*>(res->indel_bytes + e->str_off), (size_t)e->win_len);
    }
};

namespace {

// ---- GenomeRegion.smooth (:188-208), Int arithmetic ------------------------------------------
std::vector<int32_t> smooth(const int32_t* in, int64_t n, int window) {
    std::vector<int32_t> result((size_t)n, 0);
    const int half = window / 2;
    int32_t accum = 0;
    for (int64_t i = 0; i < n; i++) {
        accum = wrap32((int64_t)accum + in[i]);
        if (i > window) {
            accum = wrap32((int64_t)accum - in[i - window]);
            result[(size_t)(i - half)] = wrap32((int64_t)accum + half) / window;
        }
    }
    if (n > window) {
        for (int64_t i = 0; i < window - half; i++) result[(size_t)i] = result[(size_t)(window - half)];
        for (int64_t i = n - half; i < n; i++) result[(size_t)i] = result[(size_t)(n - half - 1)];
    } else {
        for (int64_t i = 0; i < n; i++) result[(size_t)i] = accum / (int32_t)n;
    }
    return result;
}

// ---- GenomeRegion.summaryRegions (:742-763) with nearEdge (:690) -------------------------------
template <class Test>
std::vector<std::pair<int32_t, int32_t>> summary_regions(const pb_region_out& o, Test test, int slop) {
    std::vector<std::pair<int32_t, int32_t>> regs;
    int64_t first = -1, last = -1;
    for (int64_t i = 0; i < o.size; i++) {
        if (test(i)) { last = i; if (first < 0) first = i; }
        else if (last >= 0 && i > last + slop) { regs.emplace_back((int32_t)(o.start + first), (int32_t)(o.start + last)); first = last = -1; }
    }
    if (last >= 0) regs.emplace_back((int32_t)(o.start + first), (int32_t)(o.start + last));
    std::vector<std::pair<int32_t, int32_t>> out;
    for (auto& r : regs) if (!(r.first - o.start < 100 || o.stop - r.second < 100)) out.push_back(r);
    return out;
}

std::vector<std::pair<int32_t, int32_t>> duplication_events(const pb_region_out& o) {     // :735-741
    std::vector<std::pair<int32_t, int32_t>> out;
    if (o.copy_number.empty()) return out;
    for (auto& r : summary_regions(o, [&](int64_t i) { return o.copy_number[(size_t)i] > 1; }, 2000))
        if ((int64_t)r.second + 1 - r.first > 10000) out.push_back(r);
    return out;
}

std::string region_string(const std::string& name, int64_t start, int64_t stop) {          // Region.scala:27,42
    std::string s = name + ":"; put_i64(s, start);
    if (stop + 1 - start >= 2) { s += "-"; put_i64(s, stop); }
    return s;
}

// ---- fixFixList (:557-595): stable sort by locus, of overlapping neighbours keep the larger (the first on a tie) ----
std::vector<Fix> fix_fix_list(const std::vector<Fix>& in) {
    std::vector<Fix> fixes(in);
    std::stable_sort(fixes.begin(), fixes.end(), [](const Fix& a, const Fix& b) { return a.locus < b.locus; });
    std::vector<Fix> out;
    size_t i = 0;
    if (fixes.empty()) return out;
    Fix cur = fixes[0];
    for (i = 1; i < fixes.size(); i++) {
        const Fix& nx = fixes[i];
        const int64_t c0 = cur.locus, c1 = c0 + std::max<int64_t>((int64_t)cur.was.size() - 1, 0);
        const int64_t n0 = nx.locus, n1 = n0 + std::max<int64_t>((int64_t)nx.was.size() - 1, 0);
        if (n0 <= c1 && n1 >= c0) {                                                          // Region.overlaps
            if (cur.was.size() + cur.patch.size() < nx.was.size() + nx.patch.size()) cur = nx;
        } else { out.push_back(cur); cur = nx; }
    }
    out.push_back(cur);
    return out;
}

// ---- fixIssues (:597-621).  The reference applies the de-overlapped fixes from the last to the first with
// coordinates of the unfixed region; for disjoint fixes that is one left-to-right rebuild. ----
void fix_issues(pb_region_out& o, const std::vector<Fix>& list) {
    const std::vector<Fix> fixes = fix_fix_list(list);
    if (fixes.empty()) return;
    // the source coordinates of `bases` are those of the original region only while every earlier fix kept the
    // length (the SNP pass); the second pass (indels) runs on the SNP-fixed bases, which still have them
    std::vector<uint8_t> nb;
    nb.reserve(o.bases.size() + 64);
    int64_t at = 0;                                                                          // index into o.bases
    for (const Fix& f : fixes) {
        const int64_t s = (int64_t)f.locus - o.start;
        if (s < at || s + (int64_t)f.was.size() > (int64_t)o.bases.size()) { o.stats.fix_mismatches++; continue; }
        nb.insert(nb.end(), o.bases.begin() + at, o.bases.begin() + s);
        bool same = true;
        for (size_t k = 0; k < f.was.size(); k++) same = same && upper(o.contig[(int64_t)f.locus - 1 + (int64_t)k]) == f.was[k];
        if (!same) o.stats.fix_mismatches++;                                                 // "Fix mismatch: ..." log line
        nb.insert(nb.end(), f.patch.begin(), f.patch.end());
        at = s + (int64_t)f.was.size();
    }
    nb.insert(nb.end(), o.bases.begin() + at, o.bases.end());
    o.bases.swap(nb);
}

// ---- one VCF record (Vcf.scala:74-176) ----------------------------------------------------------
struct Locus {
    int64_t c[4], q[4];
    int32_t mqSum, physCov, badPair, deletions, delQual, insertions, insQual, clips;
    uint64_t call;
};

inline Locus load_locus(const pb_region_result* r, int64_t i) {
    Locus L;
    for (int b = 0; b < 4; b++) { L.c[b] = r->base_count4[4 * i + b]; L.q[b] = r->qual_sum4[4 * i + b]; }
    L.mqSum = r->mq_sum[i]; L.physCov = r->phys_cov[i]; L.badPair = r->bad_pair[i]; L.deletions = r->deletions[i];
    L.delQual = r->del_qual[i]; L.insertions = r->insertions[i]; L.insQual = r->ins_qual[i]; L.clips = r->clips[i];
    L.call = r->call[i];
    return L;
}

void put_af(std::string& s, double af) {      // "%.2f".format(af): HALF_UP on the decimal expansion
    const double y = af * 100.0;
    long long f = (long long)std::floor(y);
    if (y - (double)f >= 0.5) f++;
    put_i64(s, f / 100); s.push_back('.'); s.push_back((char)('0' + (f / 10) % 10)); s.push_back((char)('0' + f % 10));
}

void write_record(const pb_region_out& o, std::string& out, int64_t index, bool embedded, bool indelOkArg) {
    const pb_region_result* r = o.res;
    const bool indelOk = indelOkArg && index > 0;
    const int64_t locus = (int64_t)o.start + index;
    const Locus L = load_locus(r, index);
    const int base = PB_CALL_BASE(L.call), alt = PB_CALL_ALT(L.call), kind = PB_CALL_INDEL(L.call);
    const bool homo = PB_CALL_HOMO(L.call), homoIndel = PB_CALL_HOMOINDEL(L.call);
    const int64_t score = PB_CALL_SCORE(L.call);
    const char bcBase = "ACGTN"[base], bcAlt = "ACGT"[alt];
    const bool isIns = kind == 1, isDel = kind == 2;
    const int64_t count = L.c[0] + L.c[1] + L.c[2] + L.c[3];
    const int64_t depthL = count + L.deletions;                                              // PileUp.scala:44
    const int64_t qsum = L.q[0] + L.q[1] + L.q[2] + L.q[3];
    std::string bcString;                                                                    // callString(indelOk) :169-173
    if (indelOk && (isIns || isDel)) bcString = o.indel_string(index, kind); else bcString.assign(1, bcBase);
    int o0 = base;
    if (base == 4) {                                                                         // n == 0: baseSum is still sums(order(0))
        o0 = 0;                                                                              // (every sum is zero: the stable order starts at A)
    }
    const int32_t baseDP = wrap32(L.q[o0]), altBaseDP = wrap32(L.q[alt]);
    const int32_t depth = wrap32(depthL);
    int64_t loc = locus;
    std::string rB, cB; const char* callType; int32_t refDP, altDP;
    if (indelOk && !embedded && isDel) {
        loc -= 1;
        const char rb = o.ref_base(loc);
        callType = homoIndel ? "1/1" : "0/1";
        const int32_t p = std::max(pctI(L.delQual, L.mqSum), pctI(L.deletions, wrap32((int64_t)wrap32(count) + L.deletions)));   // delPct, PileUp.scala:123
        rB.assign(1, rb); rB += bcString; cB.assign(1, rb); refDP = 100 - p; altDP = p;
    } else if (indelOk && !embedded && isIns) {
        loc -= 1;
        const char rb = o.ref_base(loc);
        callType = homoIndel ? "1/1" : "0/1";
        const int32_t p = std::max(pctI(L.insQual, L.mqSum), pctI(L.insertions, wrap32(count)));                                 // insPct, PileUp.scala:122
        rB.assign(1, rb); cB.assign(1, rb); cB += bcString; refDP = 100 - p; altDP = p;
    } else if (homo) {
        const char rb = o.ref_base(loc);
        rB.assign(1, rb); cB.assign(1, bcBase);
        if (rb == bcBase || bcString == "N") { callType = "0/0"; refDP = baseDP; altDP = altBaseDP; }
        else { callType = "1/1"; refDP = altBaseDP; altDP = baseDP; }
    } else {
        const char rb = o.ref_base(loc);
        rB.assign(1, rb); callType = "0/1";
        if (rb == bcBase) { cB.assign(1, bcAlt); refDP = baseDP; altDP = altBaseDP; }
        else { cB.assign(1, bcBase); refDP = altBaseDP; altDP = baseDP; }
    }
    const bool het = callType[0] == '0' && callType[2] == '1';
    // filters are prepended in the order LowCov, Amb, Del (:119-124): the printed order is the reverse
    std::string filter;
    if (embedded) filter += "Del";
    if (!o.cfg.diploid && het) { if (!filter.empty()) filter += ";"; filter += "Amb"; }
    if (depth < r->min_depth) { if (!filter.empty()) filter += ";"; filter += "LowCov"; }
    if (filter.empty()) filter = "PASS";
    const bool dot = (cB == "N" || cB == rB);
    const int ac = callType[0] == '0' ? (callType[2] == '0' ? 0 : 1) : 2;
    double af = 0.0;
    if (wrap32((int64_t)refDP + altDP) > 0 && !dot) af = (double)((float)altDP / (float)wrap32((int64_t)refDP + altDP));
    const int64_t meanQual = roundDivL(qsum, roundDivL((int64_t)L.mqSum * count, depthL));   // PileUp.scala:64-67
    const int64_t meanMq = roundDivL((int64_t)L.mqSum - depthL, depthL);                     // :70-72
    const int64_t qd = count > 0 ? score / count : 0;                                        // :166
    out += o.name; out.push_back('\t'); put_i64(out, loc); out += "\t.\t"; out += rB; out.push_back('\t');
    if (dot) out.push_back('.'); else out += cB;
    out.push_back('\t');
    if (indelOk && isDel) out.push_back('.'); else put_i64(out, score);
    out.push_back('\t'); out += filter; out.push_back('\t');
    out += "DP="; put_i64(out, embedded ? count : depthL);
    out += ";TD="; put_i64(out, depthL + L.badPair);
    out += ";BQ="; put_i64(out, meanQual);
    out += ";MQ="; put_i64(out, meanMq);
    out += ";QD="; put_i64(out, qd);
    out += ";BC=";
    for (int b = 0; b < 4; b++) { if (b) out.push_back(','); put_i64(out, L.c[b]); }
    if (o.cfg.vcf_qe) {
        out += ";QE=";
        for (int b = 0; b < 4; b++) { if (b) out.push_back(','); put_i64(out, L.q[b]); }
    } else {
        out += ";QP=";                                                                       // BaseSum.toStringPct, BaseSum.scala:68-71
        for (int b = 0; b < 4; b++) { if (b) out.push_back(','); put_i64(out, qsum == 0 ? 0 : (100 * L.q[b] + qsum / 2) / qsum); }
    }
    out += ";PC="; put_i64(out, L.physCov);
    out += ";IC="; put_i64(out, L.insertions);
    out += ";DC="; put_i64(out, L.deletions);
    out += ";XC="; put_i64(out, L.clips);
    out += ";AC="; put_i64(out, ac);
    out += ";AF="; put_af(out, af);
    out += "\tGT\t"; out += callType; out.push_back('\n');
    if (indelOk && kind != 0 && !embedded) write_record(o, out, index, isDel && homoIndel, false);
}

void write_dup(const pb_region_out& o, std::string& out, std::pair<int32_t, int32_t> d) {   // Vcf.scala:193-201
    const int64_t loc = (int64_t)d.first - 1;
    out += o.name; out.push_back('\t'); put_i64(out, loc); out += "\t.\t"; out.push_back(o.ref_base(loc));
    out += "\t<DUP>\t.\tPASS\tSVTYPE=DUP;SVLEN="; put_i64(out, (int64_t)d.second + 1 - d.first);
    out += ";END="; put_i64(out, d.second); out += ";IMPRECISE\tGT\t./.\n";
}

}  // namespace

extern "C" const char* pb_out_last_error(void) { return g_err_out.c_str(); }

extern "C" int pb_out_create(const pb_region_result* res, const uint8_t* contig, int64_t contig_len, const char* name,
                             int32_t start, int32_t stop, const pb_output_config* cfg, pb_region_out** out) {
    if (!res || !contig || !name || !cfg || !out) return fail_out(PB_ERR_INVALID, "null argument");
    if (start < 1 || stop < start || stop > contig_len || res->size != (int64_t)stop + 1 - start)
        return fail_out(PB_ERR_INVALID, "region does not match the result");
    const bool sparse = !res->call && res->calls;      // the call plane in sparse form: changed / ambiguous loci only
    if (!res->flags || (!res->call && !res->calls)) return fail_out(PB_ERR_INVALID, "the result must carry the flags plane and the call plane or its sparse form (calls)");
    if (sparse && res->n_calls > res->calls_cap) return fail_out(PB_ERR_INVALID, "the result's sparse call entries were truncated (calls_cap too small)");
    if (res->n_indels > res->indels_cap || res->n_indel_bytes > res->indel_bytes_cap)
        return fail_out(PB_ERR_INVALID, "the result's indel evidence was truncated (indels_cap / indel_bytes_cap too small)");
    if (res->n_indels > 0 && (!res->indels || (res->n_indel_bytes > 0 && !res->indel_bytes)))
        return fail_out(PB_ERR_INVALID, "the result must carry the indel evidence (indels, indel_bytes): insertion / deletion fixes need the strings");
    pb_region_out* o = new pb_region_out();
    o->res = res; o->contig = contig; o->contig_len = contig_len; o->name = name; o->start = start; o->stop = stop;
    o->size = res->size; o->cfg = *cfg;
    o->bases.assign(contig + (start - 1), contig + stop);                                   // originalBases / bases (:36-37)
    // ---- postProcess pass 2 (:275-283), skipped like pass 1 when the region saw no reads (:229-231) ----
    if (res->frag_coverage && res->read_count != 0) {
        double sum = 0.0;
        for (int64_t i = 0; i < o->size; i++) sum += (double)res->frag_coverage[i];          // NormalDistribution.mean
        const double baseCov = sum / (double)o->size;
        const std::vector<int32_t> sm = smooth(res->frag_coverage, o->size, 200);
        o->copy_number.resize((size_t)o->size);
        for (int64_t i = 0; i < o->size; i++)
            o->copy_number[(size_t)i] = baseCov > 0 ? (int16_t)(int64_t)std::floor((double)sm[(size_t)i] / baseCov + 0.5) : (int16_t)0;
    }
    // ---- identifyAndFixIssues (:307-380, 413) ----
    pb_out_stats& st = o->stats;
    int64_t next_call = 0;
    for (int64_t i = 0; i < o->size; i++) {
        const uint8_t fl = res->flags[i];
        if (fl & PB_FL_CONFIRMED) st.confirmed++;
        if (o->bases[(size_t)i] != 'N') st.non_n++;
        if (!(fl & (PB_FL_CHANGED | PB_FL_AMBIGUOUS))) continue;
        const int kind = (fl >> PB_FL_KIND_SHIFT) & 3;
        const int32_t loc = (int32_t)(start + i);
        uint64_t call;
        if (sparse) {
            while (next_call < res->n_calls && (int64_t)res->calls[next_call].locus_index < i) next_call++;
            if (next_call >= res->n_calls || (int64_t)res->calls[next_call].locus_index != i) {
                delete o;
                return fail_out(PB_ERR_INVALID, "sparse call entries do not match the flags plane");
            }
            call = res->calls[next_call].call;
        } else {
            call = res->call[i];
        }
        const std::string rBase(1, o->ref_base(loc)), cBase(1, "ACGTN"[PB_CALL_BASE(call)]);
        switch (kind) {
            case PB_KIND_SNP:
                if (cfg->fix_snps) o->snp_fixes.push_back(Fix{loc, rBase, cBase});
                st.snps++;
                break;
            case PB_KIND_AMB:
                if (cfg->fix_snps && !cfg->longread) {
                    if (cfg->iupac) o->small_fixes.push_back(Fix{loc, rBase, std::string(1, iupac_of(cBase[0], "ACGT"[PB_CALL_ALT(call)]))});
                    else o->snp_fixes.push_back(Fix{loc, rBase, cBase});
                    st.amb++;
                }
                break;
            case PB_KIND_INS: {
                const std::string ins = o->indel_string(i, 1);
                if (cfg->fix_indels) o->small_fixes.push_back(Fix{loc, std::string(), ins});
                st.ins++; st.ins_bases += (int64_t)ins.size();
                break;
            }
            default: {
                const std::string del = o->indel_string(i, 2);
                if (cfg->fix_indels) o->small_fixes.push_back(Fix{loc, del, std::string()});
                st.dels++; st.del_bases += (int64_t)del.size();
            }
        }
    }
    // the Scala lists are built by prepending (::=): newest first
    std::reverse(o->snp_fixes.begin(), o->snp_fixes.end());
    std::reverse(o->small_fixes.begin(), o->small_fixes.end());
    {
        char buf[160];
        snprintf(buf, sizeof buf, "Confirmed %lld of %lld bases (%.2f%%)\n", (long long)st.confirmed, (long long)st.non_n,
                 st.non_n ? (double)st.confirmed * 100.0 / (double)st.non_n : NAN);
        o->log += buf;
        o->log += cfg->fix_snps ? "Corrected " : "Found ";
        if (cfg->diploid) { put_i64(o->log, st.snps + st.amb); o->log += " snps"; }
        else { put_i64(o->log, st.snps); o->log += " snps; "; put_i64(o->log, st.amb); o->log += " ambiguous bases"; }
        o->log += cfg->fix_indels ? "; corrected " : "; found ";
        put_i64(o->log, st.ins); o->log += " small insertions totaling "; put_i64(o->log, st.ins_bases);
        o->log += " bases, "; put_i64(o->log, st.dels); o->log += " small deletions totaling "; put_i64(o->log, st.del_bases); o->log += " bases\n";
        for (auto& d : duplication_events(*o)) {
            o->log += "Large collapsed region: " + region_string(o->name, d.first, d.second) + " size ";
            put_i64(o->log, (int64_t)d.second + 1 - d.first); o->log += "\n";
            st.n_dups++;
        }
    }
    fix_issues(*o, o->snp_fixes);                                                           // :380
    fix_issues(*o, o->small_fixes);                                                         // :413 (bigFixList is empty here)
    {
        std::vector<Fix> all(o->snp_fixes); all.insert(all.end(), o->small_fixes.begin(), o->small_fixes.end());
        st.n_fixes = (int64_t)fix_fix_list(all).size();
    }
    *out = o;
    return PB_OK;
}

extern "C" int pb_out_destroy(pb_region_out* o) { delete o; return PB_OK; }

extern "C" int pb_out_stats_get(const pb_region_out* o, pb_out_stats* st) {
    if (!o || !st) return fail_out(PB_ERR_INVALID, "null argument");
    *st = o->stats; return PB_OK;
}

extern "C" int pb_out_bases(const pb_region_out* o, const uint8_t** bases, int64_t* n) {
    if (!o || !bases || !n) return fail_out(PB_ERR_INVALID, "null argument");
    *bases = o->bases.data(); *n = (int64_t)o->bases.size(); return PB_OK;
}

extern "C" int pb_out_copy_number(const pb_region_out* o, const int16_t** cn, int64_t* n) {
    if (!o || !cn || !n) return fail_out(PB_ERR_INVALID, "null argument");
    *cn = o->copy_number.data(); *n = (int64_t)o->copy_number.size(); return PB_OK;
}

extern "C" int pb_out_log(pb_region_out* o, const char** text, int64_t* n) {
    if (!o || !text || !n) return fail_out(PB_ERR_INVALID, "null argument");
    *text = o->log.data(); *n = (int64_t)o->log.size(); return PB_OK;
}

// GenomeRegion.writeChanges (:646-657)
extern "C" int pb_out_changes(pb_region_out* o, const char* new_name, int64_t offset, const char** text, int64_t* n) {
    if (!o || !text || !n) return fail_out(PB_ERR_INVALID, "null argument");
    const std::string newName = new_name ? new_name : o->name;
    std::vector<Fix> all(o->snp_fixes); all.insert(all.end(), o->small_fixes.begin(), o->small_fixes.end());
    int64_t delta = 0;
    o->changes.clear();
    for (const Fix& f : fix_fix_list(all)) {
        const int64_t loc = f.locus, newLoc = loc + delta;
        o->changes += region_string(o->name, loc, loc + (int64_t)f.was.size() - 1) + " " +
                      region_string(newName, newLoc + offset, newLoc + offset + (int64_t)f.patch.size() - 1) + " " +
                      (f.was.empty() ? "." : f.was) + " " + (f.patch.empty() ? "." : f.patch) + "\n";
        delta += (int64_t)f.patch.size() - (int64_t)f.was.size();
    }
    *text = o->changes.data(); *n = (int64_t)o->changes.size();
    return PB_OK;
}

// GenomeRegion.writeVcf (:623-643): every locus of the region, in order; spread over host threads by locus ranges
extern "C" int pb_out_vcf(pb_region_out* o, int threads, const char** text, int64_t* n) {
    if (!o || !text || !n) return fail_out(PB_ERR_INVALID, "null argument");
    const pb_region_result* r = o->res;
    if (!r->base_count4 || !r->qual_sum4 || !r->mq_sum || !r->phys_cov || !r->bad_pair || !r->deletions || !r->del_qual ||
        !r->insertions || !r->ins_qual || !r->clips || !r->call)
        return fail_out(PB_ERR_INVALID, "pb_out_vcf needs every PileUp counter plane and the call plane in the result");
    const auto dups = duplication_events(*o);
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : 1, o->size / 4096 + 1));
    std::vector<std::string> parts((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t]() {
        const int64_t i0 = o->size * t / nt, i1 = o->size * (t + 1) / nt;
        std::string& s = parts[(size_t)t];
        s.reserve((size_t)(i1 - i0) * 150);
        size_t di = 0;
        while (di < dups.size() && dups[di].first < o->start + i0) di++;
        for (int64_t i = i0; i < i1; i++) {
            if (di < dups.size() && dups[di].first == o->start + i) { write_dup(*o, s, dups[di]); di++; }
            write_record(*o, s, i, (r->flags[i] & PB_FL_DELETED) != 0, true);
        }
    });
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    o->vcf.clear(); o->vcf.reserve(total);
    for (auto& p : parts) o->vcf += p;
    *text = o->vcf.data(); *n = (int64_t)o->vcf.size();
    return PB_OK;
}

// Tracks.makeTrack body for one region (Tracks.scala:169-186): "fixedStep ..." + one value per locus
extern "C" int pb_out_wig(pb_region_out* o, int track, const char** text, int64_t* n) {
    if (!o || !text || !n) return fail_out(PB_ERR_INVALID, "null argument");
    const pb_region_result* r = o->res;
    auto need = [&](const void* p) { return p != nullptr; };
    std::string& s = o->wig;
    s.clear();
    s += "fixedStep chrom=" + o->name + " start="; put_i64(s, o->start); s += " step=1\n";
    for (int64_t i = 0; i < o->size; i++) {
        int64_t v = 0;
        switch (track) {
            case PB_TRACK_CHANGES: v = (r->flags[i] & PB_FL_CHANGED) ? 1 : 0; break;                                      // :57-60
            case PB_TRACK_UNCONFIRMED: v = (r->flags[i] & PB_FL_CONFIRMED) ? 0 : 1; break;                                // :62-65
            case PB_TRACK_COPY_NUMBER: v = (o->copy_number.empty() ? 0 : o->copy_number[(size_t)i]) - 1; break;           // :67-70
            case PB_TRACK_COVERAGE: if (!need(r->coverage_arr)) return fail_out(PB_ERR_INVALID, "plane missing"); v = r->coverage_arr[i]; break;
            case PB_TRACK_BAD_COVERAGE: if (!need(r->bad_pair)) return fail_out(PB_ERR_INVALID, "plane missing"); v = r->bad_pair[i]; break;
            case PB_TRACK_PCT_BAD: {                                                                                        // :139-147
                if (!need(r->coverage_arr) || !need(r->bad_pair)) return fail_out(PB_ERR_INVALID, "plane missing");
                const int32_t good = r->coverage_arr[i], bad = r->bad_pair[i];
                v = wrap32((int64_t)good + bad) > 0 ? wrap32((int64_t)bad * 100) / wrap32((int64_t)good + bad) : 0; break;
            }
            case PB_TRACK_PHYSICAL_COVERAGE: if (!need(r->phys_cov)) return fail_out(PB_ERR_INVALID, "plane missing"); v = r->phys_cov[i]; break;
            case PB_TRACK_CLIPPED: if (!need(r->clips)) return fail_out(PB_ERR_INVALID, "plane missing"); v = (int16_t)r->clips[i]; break;   // clips.toShort (:253)
            case PB_TRACK_WEIGHTED_QUAL: if (!need(r->weighted_qual)) return fail_out(PB_ERR_INVALID, "plane missing"); v = r->weighted_qual[i]; break;
            case PB_TRACK_WEIGHTED_MQ: if (!need(r->weighted_mq)) return fail_out(PB_ERR_INVALID, "plane missing"); v = r->weighted_mq[i]; break;
            case PB_TRACK_DELTA_COVERAGE: case PB_TRACK_DIP_COVERAGE: {                                                    // :673-688 on fragCoverage, radius 100
                if (!need(r->frag_coverage)) return fail_out(PB_ERR_INVALID, "plane missing");
                const int32_t left = r->frag_coverage[std::max<int64_t>(0, i - 100)], right = r->frag_coverage[std::min<int64_t>(o->size - 1, i + 100)];
                const int32_t center = r->frag_coverage[i];
                v = track == PB_TRACK_DELTA_COVERAGE ? (int64_t)std::abs(wrap32((int64_t)left - right))
                                                     : (int64_t)wrap32((int64_t)wrap32((int64_t)left - center) + wrap32((int64_t)right - center));
                break;
            }
            default: return fail_out(PB_ERR_INVALID, "unknown track");
        }
        put_i64(s, v); s.push_back('\n');
    }
    *text = s.data(); *n = (int64_t)s.size();
    return PB_OK;
}

// ---- GenomeFile-level helpers ----------------------------------------------------------------
static int emit(const std::string& s, char* buf, int64_t cap, int64_t* n) {
    if (n) *n = (int64_t)s.size();
    if (!buf) return PB_OK;                                          // size query
    if ((int64_t)s.size() > cap) return fail_out(PB_ERR_INVALID, "buffer too small");
    memcpy(buf, s.data(), s.size());
    return PB_OK;
}

// name + sep + "pilon" (GenomeFile.scala:137-141)
extern "C" int pb_pilon_name(const char* name, char* buf, int64_t cap, int64_t* n) {
    if (!name) return fail_out(PB_ERR_INVALID, "null argument");
    const std::string nm(name);
    const std::string sep = nm.find('|') == std::string::npos ? "_" : (nm.back() == '|' ? "" : "|");
    return emit(nm + sep + "pilon", buf, cap, n);
}

// GenomeFile.writeFastaElement (:79-82): ">" header, then the sequence 80 characters to the line
extern "C" int pb_fasta_element(const char* header, const uint8_t* bases, int64_t n_bases, char* buf, int64_t cap, int64_t* n) {
    if (!header || (!bases && n_bases)) return fail_out(PB_ERR_INVALID, "null argument");
    std::string s = ">"; s += header; s += "\n";
    s.reserve(s.size() + (size_t)n_bases + (size_t)n_bases / 80 + 2);
    for (int64_t i = 0; i < n_bases; i += 80) { s.append(reinterpret_cast<const char*>(bases + i), (size_t)std::min<int64_t>(80, n_bases - i)); s.push_back('\n'); }
    return emit(s, buf, cap, n);
}

// Vcf.writeHeader (Vcf.scala:28-68); date / version / command line / reference URI come from the driver
extern "C" int pb_vcf_header(const pb_output_config* cfg, const char* date, const char* version, const char* command_args, const char* reference_uri,
                             const char* const* contig_names, const int64_t* contig_sizes, int32_t n_contigs, char* buf, int64_t cap, int64_t* n) {
    if (!cfg || !date || !version || !command_args || !reference_uri) return fail_out(PB_ERR_INVALID, "null argument");
    std::string s;
    s += "##fileformat=VCFv4.1\n";
    s += std::string("##fileDate=") + date + "\n";
    s += std::string("##source=\"") + version + "\"\n";
    s += std::string("##PILON=\"") + command_args + "\"\n";
    s += std::string("##reference=") + reference_uri + "\n";
    for (int32_t i = 0; i < n_contigs; i++) { s += std::string("##contig=<ID=") + contig_names[i] + ",length="; put_i64(s, contig_sizes[i]); s += ">\n"; }
    s += "##FILTER=<ID=LowCov,Description=\"Low Coverage of good reads at location\">\n";
    s += "##FILTER=<ID=Amb,Description=\"Ambiguous evidence in haploid genome\">\n";
    s += "##FILTER=<ID=Del,Description=\"This base is in a deletion or change event from another record\">\n";
    s += "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Valid read depth; some reads may have been filtered\">\n";
    s += "##INFO=<ID=TD,Number=1,Type=Integer,Description=\"Total read depth including bad pairs\">\n";
    s += "##INFO=<ID=PC,Number=1,Type=Integer,Description=\"Physical coverage of valid inserts across locus\">\n";
    s += "##INFO=<ID=BQ,Number=1,Type=Integer,Description=\"Mean base quality at locus\">\n";
    s += "##INFO=<ID=MQ,Number=1,Type=Integer,Description=\"Mean read mapping quality at locus\">\n";
    s += "##INFO=<ID=QD,Number=1,Type=Integer,Description=\"Variant confidence/quality by depth\">\n";
    s += "##INFO=<ID=BC,Number=4,Type=Integer,Description=\"Count of As, Cs, Gs, Ts at locus\">\n";
    if (cfg->vcf_qe) s += "##INFO=<ID=QE,Number=4,Type=Integer,Description=\"Evidence for As, Cs, Gs, Ts weighted by Q & MQ at locus\">\n";
    else s += "##INFO=<ID=QP,Number=4,Type=Integer,Description=\"Percentage of As, Cs, Gs, Ts weighted by Q & MQ at locus\">\n";
    s += "##INFO=<ID=IC,Number=1,Type=Integer,Description=\"Number of reads with insertion here\">\n";
    s += "##INFO=<ID=DC,Number=1,Type=Integer,Description=\"Number of reads with deletion here\">\n";
    s += "##INFO=<ID=XC,Number=1,Type=Integer,Description=\"Number of reads clipped here\">\n";
    s += "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Allele count in genotypes, for each ALT allele, in the same order as listed\">\n";
    s += "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Fraction of evidence in support of alternate allele(s)\">\n";
    s += "##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">\n";
    s += "##INFO=<ID=SVLEN,Number=.,Type=String,Description=\"Difference in length between REF and ALT alleles\">\n";
    s += "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the variant described in this record\">\n";
    s += "##INFO=<ID=IMPRECISE,Number=0,Type=Flag,Description=\"Imprecise change from local reassembly (ALT contains Ns)\">\n";
    s += "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    s += "##FORMAT=<ID=AD,Number=.,Type=String,Description=\"Allelic depths for the ref and alt alleles in the order listed\">\n";
    s += "##FORMAT=<ID=DP,Number=1,Type=String,Description=\"Approximate read depth; some reads may have been filtered\">\n";
    s += "##ALT=<ID=DUP,Description=\"Possible segmental duplication\">\n";
    s += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n";
    return emit(s, buf, cap, n);
}
