/*
 * lf_chain.inl -- the chain-level operator: lordFAST's alignChain_edlib (src/LordFAST.cpp:1765-2258)
 * for a whole chunk of candidate chains at once, re-phased as collect -> batch -> emit:
 *
 *   round 1  every head (SHW, :1833), gap (NW, :1941) and tail (SHW, :2168) alignment of every chain
 *            -> one lf_gpu_align_batch-style run on the resident reads
 *   round 2  the result-dependent triggers are evaluated on the host with the reference's own float /
 *            double expressions (:1840, :1952, :2175) -> ksw_extend / ksw_extend2 tasks
 *            (:1848, :1971, :1981, :2180) -> one extend run
 *   round 3  the follow-up alignments (:1853, :2001, :2037, :2039, :2084, :2184) -> one more align run
 *   emit     per chain, the reference's op / MD accumulation (edlibCigar_push*, edlibMD_push*,
 *            :1570-1715) and string forms (edlibCigar_toString :1596, edlibMD_toString :1717),
 *            including its quirks (:2056-2057 MD/CIGAR order in the inversion branch, inclusive qEnd
 *            at :2155), producing one lf_sam_record per Sam_t the reference would push.
 *
 * Included by lf_pipeline.inl (so it is part of liblfgpu.so and of the test-only emulator build).
 */
#include <thread>

struct lf_chain_results {
    std::vector<lf_sam_record> recs;
    std::string text;
    lf_chain_stats stats;
};

namespace {

const int kClipLen = 500, kSplitLen = 80;                      /* src/LordFAST.cpp:88-92 */
const double kClipSim = 0.75, kSplitSim = 0.40, kReverseSim = 0.60;

struct SplitInfo {
    uint32_t seed_idx;     /* global index of the seed that precedes the gap */
    int32_t ext_f, ext_r;  /* extend task indices */
    int32_t t_first, t_mid_f, t_mid_r, t_second; /* round-3 task indices or -1 */
    uint32_t qs2, ts2, qe2, te2;
    bool split;
};
struct ClipInfo { int32_t ext; int32_t t3; int32_t qle, tle; };

struct ChainPlan {
    int32_t head_task = -1, tail_task = -1;   /* round-1 task indices */
    int32_t head_clip = -1, tail_clip = -1;   /* index into clips */
    uint32_t chrBeg = 0, chrEnd = 0;
    bool head_guard = false, tail_guard = false;
};

inline int pos2rid(const lf_contigs *c, int64_t pos, int64_t l_pac)
{ /* bns_pos2rid, lib/bwa/bntseq.c:349-363 */
    int left = 0, mid = 0, right = c->n;
    if (pos >= l_pac) return -1;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos >= c->offset[mid]) {
            if (mid == c->n - 1) break;
            if (pos < c->offset[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

inline lf_align_task mk_task(uint32_t rid, uint32_t qo, uint32_t ql, uint32_t to, uint32_t tl, unsigned flags, int mode)
{
    lf_align_task t;
    t.read_id = rid; t.q_off = qo; t.q_len = ql; t.t_off = to; t.t_len = tl; t.flags = (uint16_t)flags; t.mode = (uint8_t)mode; t.reserved = 0;
    return t;
}
inline lf_extend_task mk_ext(uint32_t rid, uint32_t qo, uint32_t ql, uint32_t to, uint32_t tl, unsigned flags, bool clip)
{
    lf_extend_task t;
    t.read_id = rid; t.q_off = qo; t.q_len = ql; t.t_off = to; t.t_len = tl; t.flags = (uint16_t)flags; t.matrix = LF_MAT_CLIP; t.reserved = 0;
    if (clip) { t.o_del = 0; t.e_del = 1; t.o_ins = 0; t.e_ins = 1; t.w = 40; t.zdrop = 40; }   /* :1848, :2180 */
    else { t.o_del = 8; t.e_del = 1; t.o_ins = 4; t.e_ins = 1; t.w = 100; t.zdrop = 200; }      /* :1971, :1981 */
    t.h0 = (int32_t)ql;
    return t;
}

inline char pac_base(const uint8_t *pac, uint32_t l) { return "ACGT"[(pac[l >> 2] >> ((~l & 3) << 1)) & 3]; }

/* expanded per-op buffers of one record, as the reference's two deques hold them */
struct RecBuf {
    std::string cig, md;
    void clear() { cig.clear(); md.clear(); }
    void run(char c, size_t n) { cig.append(n, c); md.append(n, c == 'I' ? '-' : '='); }
    void del_run(const uint8_t *pac, uint32_t t0, uint32_t n) { cig.append(n, 'D'); for (uint32_t k = 0; k < n; k++) md.push_back(pac_base(pac, t0 + k)); }
    /* ops of one alignment; reversed = the task ran right-to-left (pushfront in the reference);
     * t0 = forward reference position of the first target base the segment covers */
    void segment(const uint8_t *ops, const lf_align_result &r, bool reversed, const uint8_t *pac, uint32_t t0)
    {
        uint32_t tp = t0;
        for (uint32_t k = 0; k < r.ops_len; k++) {
            uint64_t p = reversed ? r.ops_off + (r.ops_len - 1 - k) : r.ops_off + k;
            unsigned op = LF_OP_AT(ops, p);
            switch (op) {
            case 0: cig.push_back('M'); md.push_back('='); tp++; break;
            case 1: cig.push_back('I'); md.push_back('-'); break;
            case 2: cig.push_back('D'); md.push_back(pac_base(pac, tp)); tp++; break;
            default: cig.push_back('M'); md.push_back(pac_base(pac, tp)); tp++; break;
            }
        }
    }
};

inline void put_num(std::string &s, long v) { char t[24]; int n = snprintf(t, sizeof t, "%ld", v); s.append(t, (size_t)n); }

void cigar_to_string(const std::string &cig, std::string &out)
{ /* edlibCigar_toString, :1596-1626: leading / trailing insert runs print as soft clips */
    char ch = 0; long num = 0; int nops = 0;
    for (size_t i = 0; i < cig.size(); i++) {
        if (cig[i] != ch) {
            if (ch != 0) { put_num(out, num); out.push_back((nops == 0 && ch == 'I') ? 'S' : ch); nops++; }
            num = 1; ch = cig[i];
        } else num++;
    }
    if (num) { put_num(out, num); out.push_back(ch == 'I' ? 'S' : ch); }
}
void md_to_string(const std::string &md, const std::string &cig, std::string &out)
{ /* edlibMD_toString, :1717-1763 */
    long num = 0; char last = '=';
    for (size_t i = 0; i < md.size(); i++) {
        char m = md[i], c = cig[i];
        if (m == '=') { num++; last = '='; }
        else if (m == '-') { last = 'I'; }
        else if (c == 'M') { put_num(out, num); num = 0; out.push_back(m); last = 'X'; }
        else if (c == 'D') { if (last != 'D') { put_num(out, num); num = 0; out.push_back('^'); } out.push_back(m); last = 'D'; }
    }
    put_num(out, num);
}

struct Emit {
    std::vector<lf_sam_record> recs;
    std::string text;
    void push(uint32_t chain_id, uint32_t flag, uint32_t pos, uint32_t posEnd, uint32_t qStart, uint32_t qEnd, int32_t nm, const RecBuf &b)
    {
        lf_sam_record r;
        r.chain_id = chain_id; r.flag = flag; r.pos = pos; r.posEnd = posEnd; r.qStart = qStart; r.qEnd = qEnd; r.nmCount = nm;
        r.cigar_off = text.size(); cigar_to_string(b.cig, text); r.cigar_len = (uint32_t)(text.size() - r.cigar_off); text.push_back('\0');
        r.md_off = text.size(); md_to_string(b.md, b.cig, text); r.md_len = (uint32_t)(text.size() - r.md_off); text.push_back('\0');
        recs.push_back(r);
    }
};

} // namespace

extern "C" {

int lf_gpu_align_chains(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_contigs *contigs, const lf_seed *seeds,
                        const lf_chain *chains, size_t n_chains, const uint8_t *pac_host, lf_chain_results **out)
{
    if (!ctx || !reads || !contigs || !seeds || (!chains && n_chains) || !pac_host || !out || contigs->n < 1) return LF_ERR_BAD_ARG;
    *out = nullptr;
    lf_chain_results *R = new lf_chain_results();
    memset(&R->stats, 0, sizeof R->stats);
    int rc;
#define LF_CH(expr) do { rc = (expr); if (rc != 0) { delete R; return rc; } } while (0)
    LF_CH(lf_gpu_upload_reads(ctx, reads));

    /* ---------------- round 1: tasks known from the chains alone ---------------- */
    std::vector<ChainPlan> plan(n_chains);
    size_t total_seeds = 0;
    for (size_t c = 0; c < n_chains; c++) { if (chains[c].n_seeds < 2 || chains[c].read_id >= reads->n_reads) { delete R; return LF_ERR_BAD_ARG; } total_seeds += chains[c].n_seeds; }
    std::vector<lf_align_task> t1;
    t1.reserve(total_seeds + 2 * n_chains);
    std::vector<int32_t> gap_task;  /* per (chain, seed i): round-1 task of the gap after seed i, or -1 */
    std::vector<uint64_t> gap_base(n_chains + 1, 0);
    gap_task.reserve(total_seeds);
    for (size_t c = 0; c < n_chains; c++) {
        const lf_chain &ch = chains[c];
        const lf_seed *s = seeds + ch.seed_off;
        const uint32_t n = ch.n_seeds;
        const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
        const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
        ChainPlan &p = plan[c];
        int rid = pos2rid(contigs, ((int64_t)s[0].tPos + (int64_t)s[n - 1].tPos) >> 1, ctx->l_pac); /* BWT.cpp:653-660 */
        if (rid < 0) { delete R; return LF_ERR_BAD_ARG; }
        p.chrBeg = (uint32_t)contigs->offset[rid];
        p.chrEnd = (uint32_t)(contigs->offset[rid] + contigs->len[rid] - 1);
        const int32_t a = (int32_t)s[0].qPos;
        if (a > 0 && (int64_t)s[0].tPos - (a + 20) >= (int64_t)p.chrBeg) {                       /* :1823-1825 */
            p.head_guard = true; p.head_task = (int32_t)t1.size();
            t1.push_back(mk_task(ch.read_id, 0, (uint32_t)a, s[0].tPos - (uint32_t)(a + 20), (uint32_t)(a + 20), strand | LF_F_REVERSE_BOTH, LF_MODE_SHW));
        }
        gap_base[c] = gap_task.size();
        for (uint32_t i = 0; i + 1 < n; i++) {
            const uint32_t qs = s[i].qPos + s[i].len, ts = s[i].tPos + s[i].len;
            const int32_t ql = (int32_t)(s[i + 1].qPos - qs), tl = (int32_t)(s[i + 1].tPos - ts);
            if (ql > 0 && tl > 0) { gap_task.push_back((int32_t)t1.size()); t1.push_back(mk_task(ch.read_id, qs, (uint32_t)ql, ts, (uint32_t)tl, strand, LF_MODE_NW)); }
            else gap_task.push_back(-1);
        }
        gap_task.push_back(-1);
        const uint32_t qs = s[n - 1].qPos + s[n - 1].len;
        const int32_t b = (int32_t)readLen - (int32_t)qs;
        if (b > 0 && s[n - 1].tPos + s[n - 1].len + (uint32_t)(b + 20) - 1 <= p.chrEnd) {          /* :2161-2163 */
            p.tail_guard = true; p.tail_task = (int32_t)t1.size();
            t1.push_back(mk_task(ch.read_id, qs, (uint32_t)b, s[n - 1].tPos + s[n - 1].len, (uint32_t)(b + 20), strand, LF_MODE_SHW));
        }
    }
    gap_base[n_chains] = gap_task.size();
    std::vector<lf_align_result> r1(t1.size());
    std::vector<uint8_t> ops1(lf_gpu_ops_capacity(t1.data(), t1.size()));
    if (!t1.empty()) {
        LF_CH(lf_gpu_upload_align_tasks(ctx, t1.data(), t1.size()));
        LF_CH(lf_gpu_run_align(ctx));
        LF_CH(lf_gpu_download_align(ctx, r1.data(), ops1.data(), ops1.size()));
    }
    R->stats.round1_tasks = t1.size();

    /* ---------------- round 2: triggers -> extensions ---------------- */
    std::vector<lf_extend_task> e2;
    std::vector<ClipInfo> clips;
    std::vector<SplitInfo> splits;
    std::vector<int32_t> gap_split(gap_task.size(), -1);
    for (size_t c = 0; c < n_chains; c++) {
        const lf_chain &ch = chains[c];
        const lf_seed *s = seeds + ch.seed_off;
        const uint32_t n = ch.n_seeds;
        const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
        ChainPlan &p = plan[c];
        if (p.head_task >= 0) {
            const lf_align_task &t = t1[(size_t)p.head_task];
            const int32_t len = (int32_t)t.q_len, ed = r1[(size_t)p.head_task].edit_distance;
            if (len > kClipLen && (1 - ((float)ed / len)) < kClipSim) {                            /* :1840 */
                p.head_clip = (int32_t)clips.size();
                clips.push_back(ClipInfo{ (int32_t)e2.size(), -1, 0, 0 });
                e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand | LF_F_REVERSE_BOTH, true));
            }
        }
        for (uint32_t i = 0; i + 1 < n; i++) {
            const int32_t gt = gap_task[gap_base[c] + i];
            if (gt < 0) continue;
            const lf_align_task &t = t1[(size_t)gt];
            const int32_t ql = (int32_t)t.q_len, tl = (int32_t)t.t_len, ed = r1[(size_t)gt].edit_distance;
            if (abs(ql - tl) >= kSplitLen && (1 - ((float)ed / ql)) < kSplitSim) {                 /* :1952 */
                gap_split[gap_base[c] + i] = (int32_t)splits.size();
                SplitInfo si; memset(&si, 0, sizeof si);
                si.seed_idx = (uint32_t)(ch.seed_off + i); si.ext_f = (int32_t)e2.size(); si.ext_r = si.ext_f + 1;
                si.t_first = si.t_mid_f = si.t_mid_r = si.t_second = -1; si.split = false;
                splits.push_back(si);
                e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand, false));
                e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand | LF_F_REVERSE_BOTH, false));
            }
        }
        if (p.tail_task >= 0) {
            const lf_align_task &t = t1[(size_t)p.tail_task];
            const int32_t len = (int32_t)t.q_len, ed = r1[(size_t)p.tail_task].edit_distance;
            if (len > kClipLen && (1 - ((float)ed / len)) < kClipSim) {                            /* :2175 */
                p.tail_clip = (int32_t)clips.size();
                clips.push_back(ClipInfo{ (int32_t)e2.size(), -1, 0, 0 });
                e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand, true));
            }
        }
    }
    std::vector<lf_extend_result> x2(e2.size());
    if (!e2.empty()) {
        LF_CH(lf_gpu_upload_extend_tasks(ctx, e2.data(), e2.size()));
        LF_CH(lf_gpu_run_extend(ctx));
        LF_CH(lf_gpu_download_extend(ctx, x2.data()));
    }
    R->stats.round2_extends = e2.size();

    /* ---------------- round 3: follow-up alignments ---------------- */
    std::vector<lf_align_task> t3;
    for (size_t c = 0; c < n_chains; c++) {
        const lf_chain &ch = chains[c];
        const lf_seed *s = seeds + ch.seed_off;
        const uint32_t n = ch.n_seeds;
        const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
        ChainPlan &p = plan[c];
        if (p.head_clip >= 0) {
            ClipInfo &ci = clips[(size_t)p.head_clip];
            const lf_align_task &t = t1[(size_t)p.head_task];
            ci.qle = x2[(size_t)ci.ext].qle; ci.tle = x2[(size_t)ci.ext].tle;
            if (ci.qle > 0 && ci.qle < (int32_t)t.q_len) {                                         /* :1850-1853 */
                ci.t3 = (int32_t)t3.size();
                t3.push_back(mk_task(ch.read_id, t.q_len - (uint32_t)ci.qle, (uint32_t)ci.qle, s[0].tPos - (uint32_t)ci.tle, (uint32_t)ci.tle,
                                     strand | LF_F_REVERSE_BOTH, LF_MODE_NW));
            }
        }
        for (uint32_t i = 0; i + 1 < n; i++) {
            const int32_t sx = gap_split[gap_base[c] + i];
            if (sx < 0) continue;
            SplitInfo &si = splits[(size_t)sx];
            const lf_align_task &t = t1[(size_t)gap_task[gap_base[c] + i]];
            const uint32_t qs = t.q_off, ts = t.t_off, qe = qs + t.q_len, te = ts + t.t_len;
            si.qs2 = qs + (uint32_t)x2[(size_t)si.ext_f].qle; si.ts2 = ts + (uint32_t)x2[(size_t)si.ext_f].tle;   /* :1972-1973 */
            si.qe2 = qe - (uint32_t)x2[(size_t)si.ext_r].qle; si.te2 = te - (uint32_t)x2[(size_t)si.ext_r].tle;   /* :1982-1983 */
            if (si.qs2 < si.qe2 || si.ts2 < si.te2) {                                              /* :1995 */
                si.split = true;
                if (si.qs2 > qs || si.ts2 > ts) {                                                  /* :1999-2001 */
                    si.t_first = (int32_t)t3.size();
                    t3.push_back(mk_task(ch.read_id, qs, si.qs2 - qs, ts, si.ts2 - ts, strand, LF_MODE_NW));
                }
                if (si.qs2 < si.qe2 && si.ts2 < si.te2) {                                          /* :2034-2039 */
                    si.t_mid_f = (int32_t)t3.size();
                    t3.push_back(mk_task(ch.read_id, si.qs2, si.qe2 - si.qs2, si.ts2, si.te2 - si.ts2, strand | LF_F_NO_PATH, LF_MODE_NW));
                    si.t_mid_r = (int32_t)t3.size();
                    t3.push_back(mk_task(ch.read_id, si.qs2, si.qe2 - si.qs2, si.ts2, si.te2 - si.ts2, strand | LF_F_RC_QUERY, LF_MODE_NW));
                }
                if (si.qe2 < qe || si.te2 < te) {                                                  /* :2080-2084 */
                    si.t_second = (int32_t)t3.size();
                    t3.push_back(mk_task(ch.read_id, si.qe2, qe - si.qe2, si.te2, te - si.te2, strand | LF_F_REVERSE_BOTH, LF_MODE_NW));
                }
            }
        }
        if (p.tail_clip >= 0) {
            ClipInfo &ci = clips[(size_t)p.tail_clip];
            const lf_align_task &t = t1[(size_t)p.tail_task];
            ci.qle = x2[(size_t)ci.ext].qle; ci.tle = x2[(size_t)ci.ext].tle;
            if (ci.qle > 0 && ci.qle < (int32_t)t.q_len) {                                         /* :2181-2184 */
                ci.t3 = (int32_t)t3.size();
                t3.push_back(mk_task(ch.read_id, t.q_off, (uint32_t)ci.qle, t.t_off, (uint32_t)ci.tle, strand, LF_MODE_NW));
            }
        }
    }
    std::vector<lf_align_result> r3(t3.size());
    std::vector<uint8_t> ops3(lf_gpu_ops_capacity(t3.data(), t3.size()));
    if (!t3.empty()) {
        LF_CH(lf_gpu_upload_align_tasks(ctx, t3.data(), t3.size()));
        LF_CH(lf_gpu_run_align(ctx));
        LF_CH(lf_gpu_download_align(ctx, r3.data(), ops3.data(), ops3.size()));
    }
    LF_CH(lf_gpu_sync(ctx));
    R->stats.round3_tasks = t3.size();
#undef LF_CH

    /* ---------------- emit: the reference's accumulation, chain by chain ---------------- */
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (n_chains < 64) nthreads = 1;
    std::vector<Emit> parts(nthreads);
    auto work = [&](unsigned tid) {
        Emit &E = parts[tid];
        RecBuf B;
        const size_t c_lo = n_chains * tid / nthreads, c_hi = n_chains * (tid + 1) / nthreads;
        for (size_t c = c_lo; c < c_hi; c++) {
            const lf_chain &ch = chains[c];
            const lf_seed *s = seeds + ch.seed_off;
            const uint32_t n = ch.n_seeds;
            const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
            const ChainPlan &p = plan[c];
            const uint32_t flag_norm = ch.is_rev ? 16u : 0u, flag_opp = ch.is_rev ? 0u : 16u;
            uint32_t flag = flag_norm, pos = s[0].tPos, qStart = s[0].qPos, posEnd = 0, qEnd = 0;
            int32_t editScore = 0;
            B.clear();
            /* head (:1820-1899) */
            const int32_t a = (int32_t)s[0].qPos;
            if (a > 0) {
                if (p.head_guard) {
                    const ClipInfo *ci = p.head_clip >= 0 ? &clips[(size_t)p.head_clip] : nullptr;
                    if (ci && ci->t3 >= 0) {
                        const lf_align_result &r = r3[(size_t)ci->t3];
                        B.run('I', (size_t)(a - ci->qle));
                        B.segment(ops3.data(), r, true, pac_host, s[0].tPos - (uint32_t)(r.end_location + 1));
                        editScore -= r.edit_distance;
                        pos = s[0].tPos - (uint32_t)r.end_location - 1;
                        qStart = s[0].qPos - (uint32_t)ci->qle;
                    } else {
                        const lf_align_result &r = r1[(size_t)p.head_task];
                        B.segment(ops1.data(), r, true, pac_host, s[0].tPos - (uint32_t)(r.end_location + 1));
                        editScore -= r.edit_distance;
                        pos = s[0].tPos - (uint32_t)r.end_location - 1;
                        qStart = 0;
                    }
                } else B.run('I', (size_t)a);
            }
            /* anchors and gaps (:1901-2137) */
            int numAnchorsSoFar = 1;
            uint32_t i = 0;
            for (; i + 1 < n; i++) {
                B.run('M', s[i].len);
                const uint32_t qs = s[i].qPos + s[i].len, ts = s[i].tPos + s[i].len;
                const uint32_t qe = s[i + 1].qPos, te = s[i + 1].tPos;
                const int32_t ql = (int32_t)(qe - qs), tl = (int32_t)(te - ts);
                if (ql > 0 && tl > 0) {
                    const int32_t gt = gap_task[gap_base[c] + i];
                    const int32_t sx = gap_split[gap_base[c] + i];
                    const SplitInfo *si = sx >= 0 ? &splits[(size_t)sx] : nullptr;
                    if (si && si->split) {
                        if (si->t_first >= 0) {
                            const lf_align_result &r = r3[(size_t)si->t_first];
                            B.segment(ops3.data(), r, false, pac_host, ts);
                            editScore -= r.edit_distance;
                        }
                        B.run('I', (size_t)(readLen - si->qs2));
                        posEnd = si->ts2; qEnd = si->qs2;
                        if (numAnchorsSoFar > 1) E.push((uint32_t)c, flag, pos, posEnd, qStart, qEnd, editScore, B);
                        B.clear(); editScore = 0;
                        if (si->t_mid_f >= 0) {
                            const lf_align_result &rf = r3[(size_t)si->t_mid_f], &rr = r3[(size_t)si->t_mid_r];
                            const int32_t ql2 = (int32_t)(si->qe2 - si->qs2);
                            if ((1 - ((double)rr.edit_distance / ql2)) > (1 - ((double)rf.edit_distance / ql2))
                                && (1 - ((double)rr.edit_distance / ql2)) > kReverseSim) {                /* :2040-2041 */
                                B.run('I', (size_t)si->qs2);
                                B.segment(ops3.data(), rr, false, pac_host, si->ts2);
                                B.cig.append((size_t)(readLen - si->qe2), 'I');
                                B.md.insert((size_t)0, (size_t)(readLen - si->qe2), '-');              /* sic, :2056-2057 */
                                E.push((uint32_t)c, flag_opp, si->ts2, si->te2, si->qs2, si->qe2, -rr.edit_distance, B);
                                B.clear();
                            }
                        }
                        B.run('I', (size_t)si->qe2);
                        if (si->t_second >= 0) {
                            const lf_align_result &r = r3[(size_t)si->t_second];
                            B.segment(ops3.data(), r, true, pac_host, si->te2);
                            editScore -= r.edit_distance;
                        }
                        flag = flag_norm; pos = si->te2; qStart = si->qe2;
                        numAnchorsSoFar = 0;
                    } else {
                        const lf_align_result &r = r1[(size_t)gt];
                        editScore -= r.edit_distance;
                        B.segment(ops1.data(), r, false, pac_host, ts);
                    }
                } else if (ql > 0) { B.run('I', (size_t)ql); editScore -= ql; }
                else { B.del_run(pac_host, ts, (uint32_t)tl); editScore -= tl; }
                numAnchorsSoFar++;
            }
            B.run('M', s[i].len);
            posEnd = s[i].tPos + s[i].len - 1;
            qEnd = s[i].qPos + s[i].len - 1;                                                       /* inclusive, :2155 */
            /* tail (:2157-2230) */
            const uint32_t qs = s[i].qPos + s[i].len;
            const int32_t b = (int32_t)readLen - (int32_t)qs;
            if (b > 0) {
                if (p.tail_guard) {
                    const uint32_t ts = s[i].tPos + s[i].len;
                    const ClipInfo *ci = p.tail_clip >= 0 ? &clips[(size_t)p.tail_clip] : nullptr;
                    if (ci && ci->t3 >= 0) {
                        const lf_align_result &r = r3[(size_t)ci->t3];
                        B.segment(ops3.data(), r, false, pac_host, ts);
                        editScore -= r.edit_distance;
                        posEnd = ts + (uint32_t)r.end_location;
                        qEnd = qs + (uint32_t)ci->qle;
                        B.run('I', (size_t)(b - ci->qle));
                    } else {
                        const lf_align_result &r = r1[(size_t)p.tail_task];
                        editScore -= r.edit_distance;
                        B.segment(ops1.data(), r, false, pac_host, ts);
                        posEnd = ts + (uint32_t)r.end_location;
                        qEnd = readLen;
                    }
                } else B.run('I', (size_t)b);
            }
            E.push((uint32_t)c, flag, pos, posEnd, qStart, qEnd, editScore, B);
        }
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    size_t nrec = 0, ntext = 0;
    for (Emit &E : parts) { nrec += E.recs.size(); ntext += E.text.size(); }
    R->recs.reserve(nrec); R->text.reserve(ntext);
    for (Emit &E : parts) {
        const uint64_t base = R->text.size();
        R->text.append(E.text);
        for (lf_sam_record r : E.recs) { r.cigar_off += base; r.md_off += base; R->recs.push_back(r); }
    }
    R->stats.records = R->recs.size();
    *out = R;
    return LF_OK;
}

const lf_sam_record *lf_chain_results_records(const lf_chain_results *r, size_t *n) { if (n) *n = r ? r->recs.size() : 0; return r ? r->recs.data() : nullptr; }
const char *lf_chain_results_text(const lf_chain_results *r, size_t *bytes) { if (bytes) *bytes = r ? r->text.size() : 0; return r ? r->text.data() : nullptr; }
int lf_chain_results_stats(const lf_chain_results *r, lf_chain_stats *out) { if (!r || !out) return LF_ERR_BAD_ARG; *out = r->stats; return LF_OK; }
void lf_chain_results_free(lf_chain_results *r) { delete r; }

} /* extern "C" */
