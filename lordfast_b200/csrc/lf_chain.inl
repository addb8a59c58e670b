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
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <time.h>

/* Result text is ~1.2 bytes per read base (240 MB per 20k x 10 kbp chunk); mapping that much fresh memory costs more
 * in page faults than filling it, so freed result buffers are parked here and reused by the next call. */
struct LfBufPool {
    const bool pinned;     /* pinned host memory (D2H target of the GPU emit) or plain malloc */
    explicit LfBufPool(bool pin) : pinned(pin) { }
    void *raw_alloc(size_t n) { return pinned ? lfb_host_alloc(n) : malloc(n); }
    void raw_free(void *p) { if (pinned) lfb_host_free(p); else free(p); }
    std::mutex mu;
    struct Item { void *p; size_t cap; };
    std::vector<Item> items;
    void *get(size_t need, size_t *cap)
    {
        std::lock_guard<std::mutex> g(mu);
        size_t bi = items.size();
        for (size_t i = 0; i < items.size(); i++)   /* best fit: the record array must not grab the text buffer */
            if (items[i].cap >= need && (bi == items.size() || items[i].cap < items[bi].cap)) bi = i;
        if (bi < items.size()) { void *p = items[bi].p; *cap = items[bi].cap; items.erase(items.begin() + (long)bi); return p; }
        if (items.size() >= 16) { raw_free(items.back().p); items.pop_back(); }
        *cap = need + need / 8 + 4096;
        return raw_alloc(*cap);
    }
    void put(void *p, size_t cap)
    {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        if (items.size() >= 16) { raw_free(p); return; }   /* room for several concurrent calls (contexts): freeing pinned memory synchronises the device */
        items.push_back(Item{p, cap});
    }
};
static LfBufPool g_result_pool(false);
static LfBufPool g_result_pool_pinned(true);

struct lf_chain_results {
    lf_sam_record *recs = nullptr; size_t n_recs = 0, recs_cap = 0;
    char *text = nullptr; size_t text_bytes = 0, text_cap = 0;
    lf_chain_stats stats;
    bool pinned = false;
    bool text_borrowed = false;   /* a lane's result: the text sits in the arena of the call that runs the lanes */
    ~lf_chain_results() { LfBufPool &P = pinned ? g_result_pool_pinned : g_result_pool; P.put(recs, recs_cap); if (!text_borrowed) P.put(text, text_cap); }
};

/* What a lane of a pipelined lf_gpu_align_chains call gets from the call: its slice of the one result text buffer
 * (records carry offsets into that buffer) and its share of the host threads. */
struct LaneIO {
    char *arena = nullptr; size_t off = 0, cap = 0;
    unsigned nthreads = 1;
};

namespace {

const int kClipLen = 500, kSplitLen = 80;                      /* src/LordFAST.cpp:88-92 */
const double kClipSim = 0.75, kSplitSim = 0.40, kReverseSim = 0.60;

struct SplitInfo {
    uint32_t seed_idx;     /* global index of the seed that precedes the gap */
    uint32_t gap_i;        /* ... and its index inside the chain */
    int32_t task;          /* the gap's round-1 task */
    int32_t ext_f, ext_r;  /* extend task indices */
    int32_t t_first, t_mid_f, t_mid_r, t_second; /* round-3 task indices or -1 */
    uint32_t qs2, ts2, qe2, te2;
    bool split;
};
struct ClipInfo { int32_t ext; int32_t t3; int32_t qle, tle; };

struct ChainPlan {
    int32_t head_task = -1, tail_task = -1;   /* round-1 task indices */
    int32_t head_clip = -1, tail_clip = -1;   /* index into clips */
    int32_t split_lo = 0, split_hi = 0;       /* the chain's entries in splits (gap order) */
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
/* Heads / tails longer than _pf_clipLen are asked for distance and end only in round 1 (k_chain_tasks does the same):
 * the reference throws their path away whenever the clip test fires and the extension shortens them (:1850-1853,
 * :2181-2184) -- every junk end -- and otherwise round 3 computes it. */
inline unsigned long_end(int32_t len) { return len > kClipLen ? (unsigned)LF_F_NO_PATH : 0u; }

inline lf_extend_task mk_ext(uint32_t rid, uint32_t qo, uint32_t ql, uint32_t to, uint32_t tl, unsigned flags, bool clip)
{
    lf_extend_task t;
    t.read_id = rid; t.q_off = qo; t.q_len = ql; t.t_off = to; t.t_len = tl; t.flags = (uint16_t)(flags & (LF_F_READ_REV | LF_F_REVERSE_BOTH)); t.matrix = LF_MAT_CLIP; t.reserved = 0;
    if (clip) { t.o_del = 0; t.e_del = 1; t.o_ins = 0; t.e_ins = 1; t.w = 40; t.zdrop = 40; }   /* :1848, :2180 */
    else { t.o_del = 8; t.e_del = 1; t.o_ins = 4; t.e_ins = 1; t.w = 100; t.zdrop = 200; }      /* :1971, :1981 */
    t.h0 = (int32_t)ql;
    return t;
}

inline char pac_base(const uint8_t *pac, uint32_t l) { return "ACGT"[(pac[l >> 2] >> ((~l & 3) << 1)) & 3]; }

inline void put_num(std::string &s, long v)
{ /* used by the rare one-char-per-op record type only */
    char t[24]; int n = 0;
    unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
    do { t[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) t[n++] = '-';
    char r[24];
    for (int k = 0; k < n; k++) r[k] = t[n - 1 - k];
    s.append(r, (size_t)n);
}
inline char *put_num(char *p, unsigned long u)
{ /* decimal digits straight into the output buffer; one and two digit numbers dominate */
    if (u < 10) { *p++ = (char)('0' + u); return p; }
    if (u < 100) { *p++ = (char)('0' + u / 10); *p++ = (char)('0' + u % 10); return p; }
    char t[24]; int n = 0;
    do { t[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (n) *p++ = t[--n];
    return p;
}

/* One record's CIGAR and MD, built as run-length strings while the pieces arrive in forward order.
 * Same output as edlibCigar_toString (:1596-1626: leading / trailing insert runs print as soft
 * clips) and edlibMD_toString (:1717-1763) applied to the reference's per-op deques, without
 * materialising one char per op: runs of matches in the 2-bit op stream are skipped a word at a time and
 * the text goes straight into per-thread buffers sized from the number of ops (<= 12 bytes per op). */
struct RecBuf {
    std::vector<char> cbuf, mbuf;
    char *cp, *mp;
    char cch; long cnum; int cnops;  /* CIGAR run in progress */
    long mnum; char mlast;           /* MD: matches since the last printed item; last move class */
    RecBuf() : cp(nullptr), mp(nullptr) { }
    void begin(size_t max_ops)
    {
        const size_t need = 12 * max_ops + 256;
        if (cbuf.size() < need) { cbuf.resize(need); mbuf.resize(need); }
        cp = cbuf.data(); mp = mbuf.data();
        cch = 0; cnum = 0; cnops = 0; mnum = 0; mlast = '=';
    }
    void clear() { cp = cbuf.data(); mp = mbuf.data(); cch = 0; cnum = 0; cnops = 0; mnum = 0; mlast = '='; }
    inline void cig_run(char c, long n)
    {
        if (n <= 0) return;
        if (c == cch) { cnum += n; return; }
        if (cch) { cp = put_num(cp, (unsigned long)cnum); *cp++ = (cnops == 0 && cch == 'I') ? 'S' : cch; cnops++; }
        cch = c; cnum = n;
    }
    inline void md_match(long n) { if (n > 0) { mnum += n; mlast = '='; } }
    inline void md_ins(long n) { if (n > 0) mlast = 'I'; }
    inline void md_mismatch(char b) { mp = put_num(mp, (unsigned long)mnum); mnum = 0; *mp++ = b; mlast = 'X'; }
    inline void md_del(char b) { if (mlast != 'D') { mp = put_num(mp, (unsigned long)mnum); mnum = 0; *mp++ = '^'; } *mp++ = b; mlast = 'D'; }
    void run(char c, size_t n) { cig_run(c, (long)n); if (c == 'I') md_ins((long)n); else md_match((long)n); }
    void del_run(const uint8_t *pac, uint32_t t0, uint32_t n) { cig_run('D', (long)n); for (uint32_t k = 0; k < n; k++) md_del(pac_base(pac, t0 + k)); }
    void finish(std::string &out_cig, std::string &out_md)
    {
        if (cnum) { cp = put_num(cp, (unsigned long)cnum); *cp++ = cch == 'I' ? 'S' : cch; }
        mp = put_num(mp, (unsigned long)mnum);
        out_cig.assign(cbuf.data(), (size_t)(cp - cbuf.data()));
        out_md.assign(mbuf.data(), (size_t)(mp - mbuf.data()));
    }
    /* ops of one alignment; reversed = the task ran right-to-left (pushfront in the reference);
     * t0 = forward reference position of the first target base the segment covers */
    void segment(const uint8_t *ops, const lf_align_result &r, bool reversed, const uint8_t *pac, uint32_t t0)
    {
        uint32_t tp = t0;
        uint32_t k = 0;
        const uint32_t n = r.ops_len;
        while (k < n) {
            /* count the run of matches starting at op k, up to 32 ops per 64-bit load */
            long run = 0;
            for (;;) {
                if (k >= n) break;
                uint64_t w; unsigned avail;
                if (!reversed) {
                    const uint64_t p = r.ops_off + k;
                    memcpy(&w, ops + (p >> 2), 8);
                    w >>= ((p & 3) << 1);
                    avail = 32 - (unsigned)(p & 3);            /* ops available in w, lowest first */
                    if (avail > n - k) avail = n - k;
                    unsigned z = w ? (unsigned)(__builtin_ctzll(w) >> 1) : 32u;
                    if (z >= avail) { run += avail; k += avail; continue; }
                    run += z; k += z;
                    break;
                } else {
                    const uint64_t p = r.ops_off + (n - 1 - k);  /* current op, walking down */
                    const uint64_t byte = p >> 2;
                    const uint64_t lo = byte >= 7 ? byte - 7 : 0;
                    memcpy(&w, ops + lo, 8);
                    const unsigned top = (unsigned)((byte - lo) * 4 + (p & 3)); /* index of op p inside w */
                    w <<= (62 - 2 * top);                        /* op p now in the two highest bits */
                    avail = top + 1;
                    if (avail > n - k) avail = n - k;
                    unsigned z = w ? (unsigned)(__builtin_clzll(w) >> 1) : 32u;
                    if (z >= avail) { run += avail; k += avail; continue; }
                    run += z; k += z;
                    break;
                }
            }
            if (run) { cig_run('M', run); md_match(run); tp += (uint32_t)run; }
            if (k >= n) break;
            const uint64_t p = reversed ? r.ops_off + (n - 1 - k) : r.ops_off + k;
            const unsigned op = LF_OP_AT(ops, p);
            k++;
            if (op == 1) { cig_run('I', 1); md_ins(1); }
            else if (op == 2) { cig_run('D', 1); md_del(pac_base(pac, tp)); tp++; }
            else { cig_run('M', 1); md_mismatch(pac_base(pac, tp)); tp++; } /* op 3 (op 0 cannot get here) */
        }
    }
};

/* The reference's accepted-inversion record pairs MD and CIGAR positions out of step (:2056-2057:
 * the trailing clip goes to the END of the CIGAR deque but to the BEGINNING of the MD deque), so that one
 * record type is built the reference's way, one char per op. */
struct SlowRec {
    std::string cig, md;
    void run(char c, size_t n) { cig.append(n, c); md.append(n, c == 'I' ? '-' : '='); }
    void segment(const uint8_t *ops, const lf_align_result &r, const uint8_t *pac, uint32_t t0)
    {
        uint32_t tp = t0;
        for (uint32_t k = 0; k < r.ops_len; k++) {
            const unsigned op = LF_OP_AT(ops, r.ops_off + k);
            switch (op) {
            case 0: cig.push_back('M'); md.push_back('='); tp++; break;
            case 1: cig.push_back('I'); md.push_back('-'); break;
            case 2: cig.push_back('D'); md.push_back(pac_base(pac, tp)); tp++; break;
            default: cig.push_back('M'); md.push_back(pac_base(pac, tp)); tp++; break;
            }
        }
    }
    void finish(std::string &out_cig, std::string &out_md)
    {
        char ch = 0; long num = 0; int nops = 0;
        for (size_t i = 0; i < cig.size(); i++) {
            if (cig[i] != ch) {
                if (ch != 0) { put_num(out_cig, num); out_cig.push_back((nops == 0 && ch == 'I') ? 'S' : ch); nops++; }
                num = 1; ch = cig[i];
            } else num++;
        }
        if (num) { put_num(out_cig, num); out_cig.push_back(ch == 'I' ? 'S' : ch); }
        long mnum = 0; char last = '=';
        for (size_t i = 0; i < md.size(); i++) {
            char m = md[i], c = cig[i];
            if (m == '=') { mnum++; last = '='; }
            else if (m == '-') { last = 'I'; }
            else if (c == 'M') { put_num(out_md, mnum); mnum = 0; out_md.push_back(m); last = 'X'; }
            else if (c == 'D') { if (last != 'D') { put_num(out_md, mnum); mnum = 0; out_md.push_back('^'); } out_md.push_back(m); last = 'D'; }
        }
        put_num(out_md, mnum);
    }
};

struct Emit {
    std::vector<lf_sam_record> recs;
    std::string tc, tm;
    RecBuf B;                 /* kept between calls together with the vectors above: no fresh pages per call */
    char *base = nullptr, *cur = nullptr, *lim = nullptr; /* this thread's slice of the result text buffer */
    bool overflow = false;
    void push_strings(uint32_t chain_id, uint32_t flag, uint32_t pos, uint32_t posEnd, uint32_t qStart, uint32_t qEnd, int32_t nm)
    {
        lf_sam_record r;
        r.chain_id = chain_id; r.flag = flag; r.pos = pos; r.posEnd = posEnd; r.qStart = qStart; r.qEnd = qEnd; r.nmCount = nm;
        if (cur + tc.size() + tm.size() + 2 > lim) { if (!overflow && getenv("LF_CHAIN_TRACE")) fprintf(stderr, "[lf_chain] text bound exceeded at chain %u: cigar %zu md %zu left %ld\n", chain_id, tc.size(), tm.size(), (long)(lim - cur)); overflow = true; return; }
        r.cigar_off = (uint64_t)(cur - base); memcpy(cur, tc.data(), tc.size()); cur += tc.size(); *cur++ = '\0'; r.cigar_len = (uint32_t)tc.size();
        r.md_off = (uint64_t)(cur - base); memcpy(cur, tm.data(), tm.size()); cur += tm.size(); *cur++ = '\0'; r.md_len = (uint32_t)tm.size();
        recs.push_back(r);
    }
    void push(uint32_t chain_id, uint32_t flag, uint32_t pos, uint32_t posEnd, uint32_t qStart, uint32_t qEnd, int32_t nm, RecBuf &b)
    {
        b.finish(tc, tm);
        push_strings(chain_id, flag, pos, posEnd, qStart, qEnd, nm);
    }
};

/* grow-only pinned host buffer kept in the context between calls */
struct PinBuf {
    void *p = nullptr; size_t cap = 0;
    void *reserve(size_t need) { if (need > cap) { lfb_host_free(p); cap = need + need / 4 + 4096; p = lfb_host_alloc(cap); if (!p) cap = 0; } return p; }
    void release() { lfb_host_free(p); p = nullptr; cap = 0; }
};
/* Device and pinned scratch of one emit list (k_emit_slots sizing pass -> scans -> writing pass). */
struct EmitSet {
    LfbBuf d_list, d_nrec, d_cigb, d_mdb, d_rec_off, d_cig_off, d_md_off, d_recs, d_text;
    LfbTemp tmp;
    PinBuf tot, h_recs;
    lfb_stream st = 0; bool have_stream = false;
    size_t n = 0, nrec = 0, ncig = 0, nmd = 0;
    void release()
    {
        LfbBuf *all[] = { &d_list, &d_nrec, &d_cigb, &d_mdb, &d_rec_off, &d_cig_off, &d_md_off, &d_recs, &d_text };
        for (LfbBuf *b : all) b->release();
        lfb_free(tmp.p); tmp.p = nullptr; tmp.cap = 0;
        tot.release(); h_recs.release();
#ifndef LF_EMU
        if (have_stream) cudaStreamDestroy(st);
#endif
        have_stream = false;
    }
};
struct LfWorkers;
struct ChainScratch {
    LfWorkers *workers = nullptr;
    PinBuf t1, r1, ops1, t3, r3, ops3, e2, x2, ed1, seeds_stage, meta_stage;
    std::vector<Emit> *parts = nullptr;
    /* device side of the GPU emit */
    EmitSet es[2];          /* [0]: chains no trigger fired for (emitted while rounds 2-3 run), [1]: the rest */
    LfbBuf d_ntask, d_contigs;
    LfbBuf d_chains, d_seeds, d_task_base, d_guards, d_clip, d_split_begin, d_splits, d_nrec, d_cigb, d_mdb, d_rec_off, d_cig_off, d_md_off, d_recs, d_text, d_ed, d_slot_base, d_slot_info, d_slot_task;
};
ChainScratch &chain_scratch(lf_gpu_ctx *ctx);

/* Host workers of one context, kept between calls: a call runs half a dozen short parallel loops, and starting 16
 * threads for each of them cost more (0.5 ms a loop) than the loops themselves. */
struct LfWorkers {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<void(unsigned)> job;
    unsigned long gen = 0;
    unsigned active = 0, pending = 0;
    bool stop = false;
    void ensure(unsigned n)
    {
        while (th.size() < n) {
            const unsigned id = (unsigned)th.size();
            th.emplace_back([this, id] {
                unsigned long seen = 0;
                for (;;) {
                    std::function<void(unsigned)> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv_go.wait(lk, [&] { return stop || (gen != seen && id < active); });
                        if (stop) return;
                        seen = gen;
                        f = job;
                    }
                    f(id);
                    { std::lock_guard<std::mutex> lk(mu); if (--pending == 0) cv_done.notify_all(); }
                }
            });
        }
    }
    void run(unsigned n, const std::function<void(unsigned)> &f)
    {   /* f(0 .. n-1), each on its own worker; returns when all are done */
        ensure(n);
        std::unique_lock<std::mutex> lk(mu);
        job = f; active = n; pending = n; gen++;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
        active = 0;
    }
    ~LfWorkers()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_go.notify_all();
        for (auto &t : th) t.join();
    }
};
static thread_local LfWorkers *tl_workers = nullptr;   /* set by lf_gpu_align_chains for the calling thread */

template <typename F>
void parallel_for(size_t n, unsigned nthreads, F fn, size_t serial_below = 256)
{ /* fn(tid, lo, hi) over contiguous ranges */
    if (nthreads <= 1 || n < serial_below) { fn(0u, (size_t)0, n); return; }
#ifndef LF_EMU
    if (tl_workers) { tl_workers->run(nthreads, [&](unsigned t) { fn(t, n * t / nthreads, n * (t + 1) / nthreads); }); return; }
#endif
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; t++) th.emplace_back(fn, t, n * t / nthreads, n * (t + 1) / nthreads);
    for (auto &t : th) t.join();
}

double now_ms()
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static double g_trace_t0 = 0;   /* LF_CHAIN_TRACE: start of the lf_gpu_align_chains call the lanes belong to */
/* LF_CHAIN_TRACE=2: GPU-side timeline.  Events recorded behind the phases of a call, printed against the reference event
 * the call recorded first (all on one device). */
#ifndef LF_EMU
struct TraceMarks {
    std::mutex mu;
    std::vector<std::pair<std::string, cudaEvent_t>> ev;
    cudaEvent_t ref = nullptr;
};
static TraceMarks g_marks;
static bool trace_gpu() { const char *e = getenv("LF_CHAIN_TRACE"); return e && atoi(e) >= 2; }
static void trace_ref(lfb_stream s)
{
    if (!trace_gpu()) return;
    std::lock_guard<std::mutex> g(g_marks.mu);
    for (auto &m : g_marks.ev) cudaEventDestroy(m.second);
    g_marks.ev.clear();
    if (!g_marks.ref) cudaEventCreate(&g_marks.ref);
    cudaEventRecord(g_marks.ref, s);
}
static void trace_mark(const lf_gpu_ctx *ctx, const char *name, lfb_stream s)
{
    if (!trace_gpu()) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s);
    char buf[96]; snprintf(buf, sizeof buf, "%p %s", (const void *)ctx, name);
    std::lock_guard<std::mutex> g(g_marks.mu);
    g_marks.ev.push_back({buf, e});
}
static void trace_print()
{
    if (!trace_gpu()) return;
    std::lock_guard<std::mutex> g(g_marks.mu);
    for (auto &m : g_marks.ev) { float ms = -1; cudaEventSynchronize(m.second); cudaEventElapsedTime(&ms, g_marks.ref, m.second); fprintf(stderr, "[gpu +%.2f] %s\n", ms, m.first.c_str()); }
}
#else
static void trace_ref(lfb_stream) { }
static void trace_mark(const lf_gpu_ctx *, const char *, lfb_stream) { }
static void trace_print() { }
#endif

void chain_scratch_free_fn(void *p)
{
    ChainScratch *s = (ChainScratch *)p;
    PinBuf *all[] = { &s->t1, &s->r1, &s->ops1, &s->t3, &s->r3, &s->ops3, &s->e2, &s->x2, &s->ed1, &s->seeds_stage, &s->meta_stage };
    for (PinBuf *b : all) b->release();
    LfbBuf *dall[] = { &s->d_ntask, &s->d_contigs, &s->d_chains, &s->d_seeds, &s->d_task_base, &s->d_guards, &s->d_clip, &s->d_split_begin, &s->d_splits, &s->d_nrec, &s->d_cigb, &s->d_mdb,
                       &s->d_rec_off, &s->d_cig_off, &s->d_md_off, &s->d_recs, &s->d_text, &s->d_ed, &s->d_slot_base, &s->d_slot_info, &s->d_slot_task };
    for (LfbBuf *b : dall) b->release();
    s->es[0].release(); s->es[1].release();
    delete s->parts;
    delete s->workers;
    delete s;
}
ChainScratch &chain_scratch(lf_gpu_ctx *ctx)
{
    if (!ctx->chain_scratch) { ctx->chain_scratch = new ChainScratch(); ctx->chain_scratch_free = chain_scratch_free_fn; }
    return *(ChainScratch *)ctx->chain_scratch;
}


/* ---- GPU emit of one list of chains: k_emit_slots sizing pass -> scans -> writing pass ---- */
int emit_size(DevState &d, EmitSet &es, LfEmitDev &E, const uint32_t *list, size_t n)
{
    lfb_stream st = es.st;
    es.n = n; es.nrec = es.ncig = es.nmd = 0;
    if (!n) return 0;
    if (es.d_list.reserve(n * 4 + 64) || es.d_nrec.reserve(n * 4 + 64) || es.d_cigb.reserve(n * 4 + 64) || es.d_mdb.reserve(n * 4 + 64)
        || es.d_rec_off.reserve((n + 1) * 8) || es.d_cig_off.reserve((n + 1) * 8) || es.d_md_off.reserve((n + 1) * 8)) return LF_ERR_NOMEM;
    unsigned long long *tot = (unsigned long long *)es.tot.reserve(64);
    if (!tot) return LF_ERR_NOMEM;
    if (h2d_k(d, es.d_list.p, list, n * 4, st, true)) return LF_ERR_CUDA;   /* `list` is pinned staging of the caller's thread (stage_get is not thread-safe) */
    E.chain_list = es.d_list.as<uint32_t>(); E.n_chains = (uint32_t)n;
    E.nrec = es.d_nrec.as<uint32_t>(); E.cig_bytes = es.d_cigb.as<uint32_t>(); E.md_bytes = es.d_mdb.as<uint32_t>();
    { auto kern = k_emit_slots<false>; LFB_LAUNCH(kern, (unsigned)n, LF_EMIT_BLOCK, 0, st, E); }   /* one block per chain, one thread per slot */
    if (lfb_scan_excl_total(es.tmp, E.nrec, es.d_rec_off.as<unsigned long long>(), n, st)
        || lfb_scan_excl_total(es.tmp, E.cig_bytes, es.d_cig_off.as<unsigned long long>(), n, st)
        || lfb_scan_excl_total(es.tmp, E.md_bytes, es.d_md_off.as<unsigned long long>(), n, st)) return LF_ERR_CUDA;
    if (lfb_d2h(&tot[0], es.d_rec_off.as<unsigned long long>() + n, 8, st) || lfb_d2h(&tot[1], es.d_cig_off.as<unsigned long long>() + n, 8, st)
        || lfb_d2h(&tot[2], es.d_md_off.as<unsigned long long>() + n, 8, st) || lfb_sync(st)) return LF_ERR_CUDA;
    es.nrec = (size_t)tot[0]; es.ncig = (size_t)tot[1]; es.nmd = (size_t)tot[2];
    return 0;
}

/* writes the list's text to text_dst (host; the bytes land at offset out_base of the caller-visible buffer) and
 * its records to es.h_recs; asynchronous on es.st */
int emit_write(EmitSet &es, LfEmitDev &E, char *text_dst, uint64_t out_base)
{
    if (!es.n) return 0;
    lfb_stream st = es.st;
    /* text_dst is pinned host memory, which the device addresses directly (unified addressing): the kernel's staged,
     * 16-byte coalesced stores go straight over PCIe while it runs, instead of to HBM and then through a copy of the
     * whole text (110 MB per config-2 chunk) after it.  LF_EMIT_NO_ZEROCOPY=1 restores the copy. */
    const bool zero_copy = !getenv("LF_EMIT_NO_ZEROCOPY");
    if (es.d_recs.reserve((es.nrec + 1) * sizeof(lf_sam_record)) || (!zero_copy && es.d_text.reserve(es.ncig + es.nmd + 64))) return LF_ERR_NOMEM;
    lf_sam_record *hr = (lf_sam_record *)es.h_recs.reserve((es.nrec + 1) * sizeof(lf_sam_record));
    if (!hr) return LF_ERR_NOMEM;
    E.rec_off = es.d_rec_off.as<uint64_t>(); E.cig_off = es.d_cig_off.as<uint64_t>(); E.md_off = es.d_md_off.as<uint64_t>();
    E.recs = es.d_recs.as<lf_sam_record>(); E.text = zero_copy ? text_dst : es.d_text.as<char>(); E.out_base = out_base;
    { auto kern = k_emit_slots<true>; LFB_LAUNCH(kern, (unsigned)es.n, LF_EMIT_BLOCK, 0, st, E); }
    if (lfb_d2h(hr, es.d_recs.p, es.nrec * sizeof(lf_sam_record), st)) return LF_ERR_CUDA;
    if (!zero_copy && lfb_d2h(text_dst, es.d_text.p, es.ncig + es.nmd, st)) return LF_ERR_CUDA;
    return 0;
}

/* the early emit (chains no trigger fired for) runs on its own host thread and stream while rounds 2-3 go on */
struct EarlyEmit {
    std::thread th;
    char *text = nullptr; size_t cap = 0;
    bool borrowed = false;
    int rc = 0;
    void join() { if (th.joinable()) th.join(); }
    ~EarlyEmit() { join(); if (text && !borrowed) g_result_pool_pinned.put(text, cap); }
};

} // namespace

extern "C" {

/* alignChain_edlib for one batch of chains on one context (the whole call, or one lane of it) */
static int align_chains_one(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_contigs *contigs, const lf_seed *seeds,
                            const lf_chain *chains, size_t n_chains, const uint8_t *pac_host, lf_chain_results **out, const LaneIO *io)
{
    if (!ctx || !reads || !contigs || !seeds || (!chains && n_chains) || !pac_host || !out || contigs->n < 1) return LF_ERR_BAD_ARG;
    *out = nullptr;
    lf_chain_results *R = new lf_chain_results();
    memset(&R->stats, 0, sizeof R->stats);
    ChainScratch &S = chain_scratch(ctx);
    if (!S.workers) S.workers = new LfWorkers();
    struct WorkersScope { LfWorkers *prev; WorkersScope(LfWorkers *w) : prev(tl_workers) { tl_workers = w; } ~WorkersScope() { tl_workers = prev; } } workers_scope(S.workers);
    unsigned nthreads = std::thread::hardware_concurrency();
    {   /* one process per GPU (torchrun): the host cores are shared by LOCAL_WORLD_SIZE of these calls; LF_HOST_THREADS overrides */
        const char *e = getenv("LF_HOST_THREADS"), *lw = getenv("LOCAL_WORLD_SIZE");
        if (e && atoi(e) > 0) nthreads = (unsigned)atoi(e);
        else if (lw && atoi(lw) > 1) nthreads = nthreads / (unsigned)atoi(lw);
    }
    if (io) nthreads = io->nthreads;
    for (DevState &dd : ctx->devs) stage_reset(dd);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    int rc;
    const double tm0 = now_ms();
#define LF_CH(expr) do { rc = (expr); if (rc != 0) { delete R; return rc; } } while (0)
    /* CIGAR / MD assembly runs on the GPU (k_emit_slots) when the context drives one device; a context over
     * several devices assembles on host threads from the 2-bit op stream (LF_CHAIN_HOST_EMIT=1 forces that). */
    const bool gpu_emit = ctx->devs.size() == 1 && !getenv("LF_CHAIN_HOST_EMIT");
    if (!io) trace_ref(ctx->devs[0].stream);
    if (reads->bases) LF_CH(lf_gpu_upload_reads(ctx, reads)); /* asynchronous: overlaps the task generation below */
    else if (io || ctx->devs.size() != 1 || ctx->devs[0].n_reads != reads->n_reads || !ctx->devs[0].n_reads) { delete R; return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_align_chains: bases == NULL needs the same reads resident on a single-device context"); }
    /* (bases == NULL: the reads lf_gpu_seed_batch / lf_gpu_upload_reads left on the device, bit planes included, are used as they are) */
    trace_mark(ctx, "reads packed", ctx->devs[0].stream);
    const double tt1 = now_ms();
    /* single-device contexts: the task list is generated on the device (k_chain_tasks); the host only needs the few
     * tasks a trigger fires for, and re-derives those from the chain (task_at below) */
    const bool dev_tasks = gpu_emit && !getenv("LF_CHAIN_HOST_TASKS");
    /* ... and so is pass A (k_chain_plan): boundaries, guards, task counts, trigger candidates, validation.  What comes back
     * is 9 B per chain.  LF_CHAIN_HOST_PLAN=1: the host loop instead (also the path of the multi-device host emit).  The
     * seeds, the plan kernel and its answer go through the extension stream (idle until round 1 is under way): on the
     * main stream the host would wait for the 200 MB of reads queued before them. */
    const bool dev_plan = dev_tasks && n_chains > 0 && !getenv("LF_CHAIN_HOST_PLAN");
    lfb_stream ps = dev_plan ? ctx->devs[0].ext_stream : ctx->devs[0].stream;
#ifndef LF_EMU
    if (dev_plan) cudaStreamWaitEvent(ps, ctx->devs[0].off_ev, 0);
#endif
    if (gpu_emit) {
        DevState &d = ctx->devs[0];
        size_t ns = 0;
        for (size_t c = 0; c < n_chains; c++) { const size_t e = (size_t)chains[c].seed_off + chains[c].n_seeds; if (e > ns) ns = e; }
        if (S.d_chains.reserve(n_chains * sizeof(lf_chain) + 64) || S.d_seeds.reserve(ns * sizeof(lf_seed) + 64)) { delete R; return LF_ERR_NOMEM; }
        /* A seed array in ordinary (pageable) memory is staged through pinned memory with all host threads, so that the copy
         * to the device is asynchronous; a caller that keeps its seeds and chains in pinned memory (lf_gpu_host_alloc) saves
         * that pass over 12 B per seed.  Either way the device reads them by kernel, not by the copy engine: there the copy
         * would wait behind the reads. */
        const size_t sb = ns * sizeof(lf_seed), cb = (sb + 63) & ~(size_t)63;
        const char *src_seeds = (const char *)seeds, *src_chains = (const char *)chains;
        if (!(lfb_is_pinned(seeds) && !((uintptr_t)seeds & 15u)) || !(lfb_is_pinned(chains) && !((uintptr_t)chains & 15u))) {
            char *stage = (char *)S.seeds_stage.reserve(ns * sizeof(lf_seed) + n_chains * sizeof(lf_chain) + 128);
            if (!stage) { delete R; return LF_ERR_NOMEM; }
            parallel_for(sb, nthreads, [&](unsigned, size_t lo, size_t hi) { memcpy(stage + lo, (const char *)seeds + lo, hi - lo); }, 1 << 20);
            memcpy(stage + cb, chains, n_chains * sizeof(lf_chain));
            src_seeds = stage; src_chains = stage + cb;
        }
        if (h2d_k(d, S.d_seeds.p, src_seeds, sb, ps, true) || h2d_k(d, S.d_chains.p, src_chains, n_chains * sizeof(lf_chain), ps, true)) { delete R; return LF_ERR_CUDA; }
    }

    /* ---------------- round 1: tasks known from the chains alone ---------------- */
    std::vector<ChainPlan> plan(n_chains);
    std::vector<uint64_t> gap_base(n_chains + 1, 0), task_base(n_chains + 1, 0);
    for (size_t c = 0; c < n_chains; c++) {
        if (chains[c].n_seeds < 2 || chains[c].read_id >= reads->n_reads) { delete R; return LF_ERR_BAD_ARG; }
        gap_base[c + 1] = gap_base[c] + chains[c].n_seeds;
    }
    const size_t total_seeds = gap_base[n_chains];
    /* per (chain, seed i): round-1 task of the gap after seed i, or -1 -- only the host emit walks these */
    std::vector<int32_t> gap_task(gpu_emit ? 0 : total_seeds, -1);
    std::vector<uint32_t> ntask(n_chains, 0);
    std::vector<uint8_t> hascand(n_chains, 0);   /* the chain holds a task whose lengths qualify for a clip / split trigger */
    std::vector<uint64_t> nslot(n_chains, 0); /* op-slot words of the chain's round-1 tasks */
    std::atomic<bool> bad_rid(false);   /* a chain the reference's chaining could not have produced (or a contig lookup that failed) */
    const double tt2 = now_ms();
    if (dev_plan) {
        DevState &d = ctx->devs[0];
        const size_t nc = (size_t)contigs->n;
        if (S.d_task_base.reserve((n_chains + 1) * 8 + 64) || S.d_guards.reserve(n_chains + 64) || S.d_ntask.reserve(n_chains * 4 + 64)
            || S.d_contigs.reserve(nc * 12 + 128)) { delete R; return LF_ERR_NOMEM; }
        const size_t co_b = (nc * 8 + 63) & ~(size_t)63;
        char *cs = (char *)stage_get(d, co_b + nc * 4 + 64);
        const size_t back_b = (n_chains + 1) * 8 + n_chains + 64;
        char *back = (char *)S.r1.reserve(back_b);   /* pinned landing zone: task_base, guards, bad flag */
        if (!cs || !back) { delete R; return LF_ERR_NOMEM; }
        memcpy(cs, contigs->offset, nc * 8); memcpy(cs + co_b, contigs->len, nc * 4);
        uint32_t *d_bad = d.queue.as<uint32_t>();   /* the work-counter block is idle until run_align clears it */
        if (d.queue.reserve(256) || h2d_k(d, S.d_contigs.p, cs, co_b + nc * 4, ps, true) || lfb_memset((d_bad = d.queue.as<uint32_t>()) + 60, 0, 4, ps)) { delete R; return LF_ERR_CUDA; }
        LFB_LAUNCH(k_chain_plan, (unsigned)((n_chains + 3) / 4), 128, 0, ps, S.d_chains.as<lf_chain>(), S.d_seeds.as<lf_seed>(), d.read_off.as<uint64_t>(), reads->n_reads,
                   (const int64_t *)S.d_contigs.p, (const int32_t *)((const char *)S.d_contigs.p + co_b), (int)nc, ctx->l_pac, (uint32_t)n_chains,
                   S.d_guards.as<uint8_t>(), S.d_ntask.as<uint32_t>(), d_bad + 60);
        if (lfb_scan_excl_total(d.tmp, S.d_ntask.as<uint32_t>(), S.d_task_base.as<unsigned long long>(), n_chains, ps)
            || lfb_d2h(back, S.d_task_base.p, (n_chains + 1) * 8, ps) || lfb_d2h(back + (n_chains + 1) * 8, S.d_guards.p, n_chains, ps)
            || lfb_d2h(back + (n_chains + 1) * 8 + n_chains + (8 - n_chains % 8) % 8, d_bad + 60, 4, ps) || lfb_sync(ps)) { delete R; return LF_ERR_CUDA; }
#ifndef LF_EMU
        cudaEventRecord(d.ext_ev, ps);               /* what the main stream launches next reads the seeds and the plan */
        cudaStreamWaitEvent(d.stream, d.ext_ev, 0);
#endif
        uint32_t badflag; memcpy(&badflag, back + (n_chains + 1) * 8 + n_chains + (8 - n_chains % 8) % 8, 4);
        if (badflag) bad_rid = true;
        memcpy(task_base.data(), back, (n_chains + 1) * 8);
        const uint8_t *gd = (const uint8_t *)back + (n_chains + 1) * 8;
        for (size_t c = 0; c < n_chains; c++) {
            plan[c].head_guard = (gd[c] & 1) != 0; plan[c].tail_guard = (gd[c] & 2) != 0; hascand[c] = (gd[c] >> 2) & 1;
            ntask[c] = (uint32_t)(task_base[c + 1] - task_base[c]);
        }
    } else
    /* pass A: boundaries, guards and task counts per chain */
    parallel_for(n_chains, nthreads, [&](unsigned, size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; c++) {
            const lf_chain &ch = chains[c];
            const lf_seed *s = seeds + ch.seed_off;
            const uint32_t n = ch.n_seeds;
            const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
            ChainPlan &p = plan[c];
            int rid = pos2rid(contigs, ((int64_t)s[0].tPos + (int64_t)s[n - 1].tPos) >> 1, ctx->l_pac); /* BWT.cpp:653-660 */
            if (rid < 0) { bad_rid = true; continue; }
            {   /* seeds inside the read and the reference, in order, without overlap (Chain.cpp:258-266 guarantees it) */
                bool okc = true;
                for (uint32_t i = 0; i < n && okc; i++) {
                    okc = s[i].len >= 1 && (uint64_t)s[i].qPos + s[i].len <= readLen && (int64_t)s[i].tPos + s[i].len <= ctx->l_pac;
                    if (okc && i + 1 < n) okc = s[i + 1].qPos >= s[i].qPos + s[i].len && s[i + 1].tPos >= s[i].tPos + s[i].len;
                }
                if (!okc) { bad_rid = true; continue; }
            }
            p.chrBeg = (uint32_t)contigs->offset[rid];
            p.chrEnd = (uint32_t)(contigs->offset[rid] + contigs->len[rid] - 1);
            uint32_t cnt = 0;
            uint64_t slots = 0;
            const int32_t a = (int32_t)s[0].qPos;
            p.head_guard = a > 0 && (int64_t)s[0].tPos - (a + 20) >= (int64_t)p.chrBeg;                       /* :1823-1825 */
            bool cand = false;
            if (p.head_guard) { cnt++; if (a <= kClipLen) slots += ((uint64_t)(2 * a + 20) + 15) >> 4; cand |= a > kClipLen; }
            for (uint32_t i = 0; i + 1 < n; i++) {
                const uint32_t qs = s[i].qPos + s[i].len, ts = s[i].tPos + s[i].len;
                const int32_t ql = (int32_t)(s[i + 1].qPos - qs), tl = (int32_t)(s[i + 1].tPos - ts);
                if (ql > 0 && tl > 0) { cnt++; slots += ((uint64_t)ql + (uint64_t)tl + 15) >> 4; cand |= abs(ql - tl) >= kSplitLen; }
            }
            const uint32_t qs = s[n - 1].qPos + s[n - 1].len;
            const int32_t b = (int32_t)readLen - (int32_t)qs;
            p.tail_guard = b > 0 && s[n - 1].tPos + s[n - 1].len + (uint32_t)(b + 20) - 1 <= p.chrEnd;         /* :2161-2163 */
            if (p.tail_guard) { cnt++; if (b <= kClipLen) slots += ((uint64_t)(2 * b + 20) + 15) >> 4; cand |= b > kClipLen; }
            ntask[c] = cnt; nslot[c] = slots; hascand[c] = cand ? 1 : 0;
        }
    });
    const double tt3 = now_ms();
    if (bad_rid) { delete R; return fail(ctx, LF_ERR_BAD_ARG, "a chain is not one the reference's chaining can produce (seeds out of order, overlapping, or outside the read / reference)"); }
    if (!dev_plan) for (size_t c = 0; c < n_chains; c++) task_base[c + 1] = task_base[c] + ntask[c];
    const size_t n1 = task_base[n_chains];
    lf_align_task *t1 = dev_tasks ? nullptr : (lf_align_task *)S.t1.reserve((n1 + 1) * sizeof(lf_align_task));
    lf_align_result *r1 = gpu_emit ? nullptr : (lf_align_result *)S.r1.reserve((n1 + 1) * sizeof(lf_align_result));
    if ((!dev_tasks && !t1) || (!gpu_emit && !r1)) { delete R; return LF_ERR_NOMEM; }
    std::vector<uint32_t> trig;   /* round-1 tasks whose result fires a clip / split trigger, ascending */
    const bool spec = gpu_emit && !getenv("LF_CHAIN_NO_SPEC");
    std::vector<uint32_t> cand_ti, cand_ext;   /* speculative round 2: candidate round-1 tasks (ascending) and their first extension */
    std::vector<lf_extend_task> se2;
    lf_extend_result *sx2 = nullptr;
    /* an early return must not leave the speculative extensions (and their D2H into pinned staging) in flight */
    struct ExtGuard { lf_gpu_ctx *c; ~ExtGuard() { if (c) { DevState &d = c->devs[0]; if (!set_dev(d)) lfb_sync(d.ext_stream); } } } ext_guard{nullptr};
    const double tt4 = now_ms();
    /* pass B: fill */
    parallel_for(n_chains, nthreads, [&](unsigned, size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; c++) {
            const lf_chain &ch = chains[c];
            const lf_seed *s = seeds + ch.seed_off;
            const uint32_t n = ch.n_seeds;
            const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
            const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
            ChainPlan &p = plan[c];
            size_t k = task_base[c];
            const int32_t a = (int32_t)s[0].qPos;
            if (dev_tasks) {   /* only the positions of the head and tail tasks */
                if (p.head_guard) p.head_task = (int32_t)task_base[c];
                if (p.tail_guard) p.tail_task = (int32_t)task_base[c + 1] - 1;
                continue;
            }
            if (p.head_guard) {
                p.head_task = (int32_t)k;
                t1[k++] = mk_task(ch.read_id, 0, (uint32_t)a, s[0].tPos - (uint32_t)(a + 20), (uint32_t)(a + 20), strand | LF_F_REVERSE_BOTH | long_end(a), LF_MODE_SHW);
            }
            for (uint32_t i = 0; i + 1 < n; i++) {
                const uint32_t qs = s[i].qPos + s[i].len, ts = s[i].tPos + s[i].len;
                const int32_t ql = (int32_t)(s[i + 1].qPos - qs), tl = (int32_t)(s[i + 1].tPos - ts);
                if (ql > 0 && tl > 0) { if (!gpu_emit) gap_task[gap_base[c] + i] = (int32_t)k; t1[k++] = mk_task(ch.read_id, qs, (uint32_t)ql, ts, (uint32_t)tl, strand, LF_MODE_NW); }
            }
            if (p.tail_guard) {
                const uint32_t qs = s[n - 1].qPos + s[n - 1].len;
                const int32_t b = (int32_t)readLen - (int32_t)qs;
                p.tail_task = (int32_t)k;
                t1[k++] = mk_task(ch.read_id, qs, (uint32_t)b, s[n - 1].tPos + s[n - 1].len, (uint32_t)(b + 20), strand | long_end(b), LF_MODE_SHW);
            }
        }
    });
    const double tt5 = now_ms();
    /* round-1 task ti of chain c, from the chain itself (the order of pass B / k_chain_tasks: head, gaps, tail) */
    auto task_at = [&](size_t c, size_t ti) -> lf_align_task {
        if (!dev_tasks) return t1[ti];
        const lf_chain &ch = chains[c];
        const lf_seed *s = seeds + ch.seed_off;
        const uint32_t n = ch.n_seeds;
        const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
        const ChainPlan &p = plan[c];
        size_t k = task_base[c];
        if (p.head_guard) {
            const int32_t a = (int32_t)s[0].qPos;
            if (k == ti) return mk_task(ch.read_id, 0, (uint32_t)a, s[0].tPos - (uint32_t)(a + 20), (uint32_t)(a + 20), strand | LF_F_REVERSE_BOTH | long_end(a), LF_MODE_SHW);
            k++;
        }
        for (uint32_t i = 0; i + 1 < n; i++) {
            const uint32_t qs = s[i].qPos + s[i].len, ts = s[i].tPos + s[i].len;
            const int32_t ql = (int32_t)(s[i + 1].qPos - qs), tl = (int32_t)(s[i + 1].tPos - ts);
            if (ql > 0 && tl > 0) { if (k == ti) return mk_task(ch.read_id, qs, (uint32_t)ql, ts, (uint32_t)tl, strand, LF_MODE_NW); k++; }
        }
        const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
        const uint32_t qs = s[n - 1].qPos + s[n - 1].len;
        const int32_t b = (int32_t)readLen - (int32_t)qs;
        return mk_task(ch.read_id, qs, (uint32_t)b, s[n - 1].tPos + s[n - 1].len, (uint32_t)(b + 20), strand | long_end(b), LF_MODE_SHW);
    };
    /* per-chain inputs of the GPU emit that are known now */
    std::vector<uint64_t> slot_base;   /* a chain of n anchors has n + 1 slots: head, n - 1 gaps, tail */
    std::vector<uint8_t> guards;
    size_t n_slots = 0;
    if (gpu_emit) {
        DevState &d = ctx->devs[0];
        slot_base.resize(n_chains + 1); guards.resize(n_chains + 1);
        for (size_t c = 0; c <= n_chains; c++) slot_base[c] = gap_base[c] + c;
        for (size_t c = 0; c < n_chains; c++) guards[c] = (uint8_t)((plan[c].head_guard ? 1 : 0) | (plan[c].tail_guard ? 2 : 0));
        n_slots = (size_t)slot_base[n_chains];
        if (S.d_task_base.reserve((n_chains + 1) * 8 + 64) || S.d_slot_base.reserve((n_chains + 1) * 8 + 64) || S.d_guards.reserve(n_chains + 64)
            || S.d_slot_info.reserve((n_slots + 1) * sizeof(LfSlotInfo)) || S.d_slot_task.reserve((n_slots + 1) * 4)) { delete R; return LF_ERR_NOMEM; }
        /* through pinned memory: a copy from pageable memory blocks the host until the stream (busy with the reads) gets to it */
        const size_t mb = ((n_chains + 1) * 8 + 63) & ~(size_t)63;
        char *ms = (char *)S.meta_stage.reserve(2 * mb + n_chains + 128);
        if (!ms) { delete R; return LF_ERR_NOMEM; }
        memcpy(ms, task_base.data(), (n_chains + 1) * 8); memcpy(ms + mb, slot_base.data(), (n_chains + 1) * 8); memcpy(ms + 2 * mb, guards.data(), n_chains);
        if (h2d_k(d, S.d_slot_base.p, ms + mb, (n_chains + 1) * 8, d.stream, true)) { delete R; return LF_ERR_CUDA; }
        if (!dev_plan && (h2d_k(d, S.d_task_base.p, ms, (n_chains + 1) * 8, d.stream, true) || h2d_k(d, S.d_guards.p, ms + 2 * mb, n_chains, d.stream, true))) { delete R; return LF_ERR_CUDA; }   /* dev_plan: k_chain_plan left them there */
    }
    size_t cap1 = 64;
    for (size_t c = 0; c < n_chains; c++) cap1 += nslot[c] * 4;
    uint8_t *ops1 = gpu_emit ? nullptr : (uint8_t *)S.ops1.reserve(cap1 + 64);
    if (!gpu_emit && !ops1) { delete R; return LF_ERR_NOMEM; }
    const double tm1 = now_ms();
    if (getenv("LF_CHAIN_TRACE")) fprintf(stderr, "[lf_chain %p +%.2f] ", (void *)ctx, now_ms() - g_trace_t0), fprintf(stderr, "tasks: upload_reads call %.2f, seed staging %.2f, pass A %.2f, prefix %.2f, pass B %.2f, emit inputs %.2f\n", tt1 - tm0, tt2 - tt1, tt3 - tt2, tt4 - tt3, tt5 - tt4, tm1 - tt5);
    if (n1) {
        if (dev_tasks) {
            DevState &d = ctx->devs[0];
            lf_align_task *dt = resident_tasks_alloc(ctx, n1);
            if (!dt) { delete R; return LF_ERR_NOMEM; }
            LFB_LAUNCH(k_chain_tasks, (unsigned)((n_chains + 3) / 4), 128, 0, d.stream, S.d_chains.as<lf_chain>(), S.d_seeds.as<lf_seed>(), d.read_off.as<uint64_t>(),
                       S.d_task_base.as<uint64_t>(), S.d_guards.as<uint8_t>(), (uint32_t)n_chains, dt);
        } else LF_CH(lf_gpu_upload_align_tasks(ctx, t1, n1));
        const double tr0 = now_ms();
        LF_CH(lf_gpu_run_align(ctx));   /* returns once the class kernels are launched (its class-count sync has waited for the uploads) */
        for (DevState &dd : ctx->devs) if (dd.cls_count[LF_CLS_BAD]) { delete R; return fail(ctx, LF_ERR_BAD_ARG, "a chain yields an alignment task outside its read or the reference"); }
        trace_mark(ctx, "round-1 kernels done", ctx->devs[0].stream);
        if (spec) {
            /* Speculative round 2: the clip / split triggers can only fire for tasks whose LENGTHS qualify (:1840 / :2175
             * ql > 500, :1952 |ql - tl| >= 80), which is known now.  Their extensions (a superset of what round 2 will
             * ask for, and nearly the same set) run on their own high-priority stream beside the round-1 kernels, so
             * round 2 costs no GPU round trip. */
            std::vector<lf_align_task> cand_task;
            if (dev_tasks) {   /* pass A has marked the chains that hold a candidate: walk those */
                for (size_t c = 0; c < n_chains; c++) {
                    if (!hascand[c]) continue;
                    /* one walk over the chain in task order (head, gaps, tail), as in task_at */
                    const lf_chain &ch = chains[c];
                    const lf_seed *s = seeds + ch.seed_off;
                    const uint32_t n = ch.n_seeds;
                    const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
                    const ChainPlan &p = plan[c];
                    size_t k = task_base[c];
                    if (p.head_guard) {
                        const int32_t a = (int32_t)s[0].qPos;
                        if (a > kClipLen) { cand_ti.push_back((uint32_t)k); cand_task.push_back(mk_task(ch.read_id, 0, (uint32_t)a, s[0].tPos - (uint32_t)(a + 20), (uint32_t)(a + 20), strand | LF_F_REVERSE_BOTH, LF_MODE_SHW)); }
                        k++;
                    }
                    for (uint32_t i = 0; i + 1 < n; i++) {
                        const uint32_t qs = s[i].qPos + s[i].len, ts = s[i].tPos + s[i].len;
                        const int32_t ql = (int32_t)(s[i + 1].qPos - qs), tl = (int32_t)(s[i + 1].tPos - ts);
                        if (ql > 0 && tl > 0) {
                            if (abs(ql - tl) >= kSplitLen) { cand_ti.push_back((uint32_t)k); cand_task.push_back(mk_task(ch.read_id, qs, (uint32_t)ql, ts, (uint32_t)tl, strand, LF_MODE_NW)); }
                            k++;
                        }
                    }
                    if (p.tail_guard) {
                        const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
                        const uint32_t qs = s[n - 1].qPos + s[n - 1].len;
                        const int32_t b = (int32_t)readLen - (int32_t)qs;
                        if (b > kClipLen) { cand_ti.push_back((uint32_t)k); cand_task.push_back(mk_task(ch.read_id, qs, (uint32_t)b, s[n - 1].tPos + s[n - 1].len, (uint32_t)(b + 20), strand, LF_MODE_SHW)); }
                    }
                }
            } else {
                std::vector<std::vector<uint32_t>> part(nthreads);
                parallel_for(n1, nthreads, [&](unsigned tid, size_t lo, size_t hi) {
                    for (size_t i = lo; i < hi; i++) {
                        const lf_align_task &t = t1[i];
                        const int32_t ql = (int32_t)t.q_len, tl = (int32_t)t.t_len;
                        if (t.mode == LF_MODE_SHW ? ql > kClipLen : abs(ql - tl) >= kSplitLen) part[tid].push_back((uint32_t)i);
                    }
                });
                for (auto &v : part) cand_ti.insert(cand_ti.end(), v.begin(), v.end());   /* ascending: the parts are consecutive ranges */
                for (const uint32_t ti : cand_ti) cand_task.push_back(t1[ti]);
            }
            cand_ext.reserve(cand_ti.size());
            for (size_t x = 0; x < cand_ti.size(); x++) {
                const lf_align_task &t = cand_task[x];
                cand_ext.push_back((uint32_t)se2.size());
                if (t.mode == LF_MODE_SHW) se2.push_back(mk_ext(t.read_id, t.q_off, t.q_len, t.t_off, t.t_len, t.flags, true));   /* head tasks carry REVERSE_BOTH already */
                else {
                    const unsigned strand = t.flags & LF_F_READ_REV;
                    se2.push_back(mk_ext(t.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand, false));
                    se2.push_back(mk_ext(t.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand | LF_F_REVERSE_BOTH, false));
                }
            }
            if (!se2.empty()) {
                lf_extend_task *pe = (lf_extend_task *)S.e2.reserve(se2.size() * sizeof(lf_extend_task));
                sx2 = (lf_extend_result *)S.x2.reserve((se2.size() + 1) * sizeof(lf_extend_result));
                if (!pe || !sx2) { delete R; return LF_ERR_NOMEM; }
                memcpy(pe, se2.data(), se2.size() * sizeof(lf_extend_task));
                ext_guard.c = ctx;
                LF_CH(spec_extend_start(ctx, pe, se2.size(), sx2));
                trace_mark(ctx, "speculative extensions done", ctx->devs[0].ext_stream);
            }
        }
        if (getenv("LF_CHAIN_TRACE") && atoi(getenv("LF_CHAIN_TRACE")) == 1) { const double tr1 = now_ms(); lf_gpu_sync(ctx); fprintf(stderr, "[lf_chain %p +%.2f] ", (void *)ctx, now_ms() - g_trace_t0), fprintf(stderr, "r1: enqueue uploads %.2f, run_align returns after %.2f (its class-count sync waits for the uploads), kernels drained after %.2f more\n", tr0 - tm1, tr1 - tr0, now_ms() - tr1); }
        if (!gpu_emit) LF_CH(lf_gpu_download_align(ctx, r1, ops1, cap1));
        else { /* results and ops stay in HBM; the trigger tests run there and only the hits come back */
            DevState &d = ctx->devs[0];
            const uint32_t cap = 1u << 16;
            if (S.d_ed.reserve(((size_t)cap + 1) * 4)) { delete R; return LF_ERR_NOMEM; }
            uint32_t *hl = (uint32_t *)S.ed1.reserve(((size_t)cap + 1) * 4);
            if (!hl) { delete R; return LF_ERR_NOMEM; }
            uint32_t *dl = S.d_ed.as<uint32_t>();   /* [0]: count, [1..]: task indices */
            if (lfb_memset(dl, 0, 4, d.stream)) { delete R; return LF_ERR_CUDA; }
            LFB_LAUNCH(k_chain_triggers, (unsigned)((n1 + 255) / 256), 256, 0, d.stream, d.tasks.as<lf_align_task>(), d.res.as<lf_align_result>(), (uint32_t)n1, dl + 1, dl, cap);
            if (lfb_d2h(hl, dl, 4, d.stream) || lfb_sync(d.stream)) { delete R; return LF_ERR_CUDA; }
            size_t nt = hl[0];
            if (nt > cap) {   /* more hits than the list holds: run again with room for all */
                if (S.d_ed.reserve((nt + 1) * 4) || !(hl = (uint32_t *)S.ed1.reserve((nt + 1) * 4))) { delete R; return LF_ERR_NOMEM; }
                dl = S.d_ed.as<uint32_t>();
                if (lfb_memset(dl, 0, 4, d.stream)) { delete R; return LF_ERR_CUDA; }
                LFB_LAUNCH(k_chain_triggers, (unsigned)((n1 + 255) / 256), 256, 0, d.stream, d.tasks.as<lf_align_task>(), d.res.as<lf_align_result>(), (uint32_t)n1, dl + 1, dl, (uint32_t)nt);
            }
            if (nt && (lfb_d2h(hl + 1, dl + 1, nt * 4, d.stream) || lfb_sync(d.stream))) { delete R; return LF_ERR_CUDA; }
            trig.assign(hl + 1, hl + 1 + nt);
            std::sort(trig.begin(), trig.end());
        }
    }
    if (!gpu_emit && n1) {   /* host copy of the results: the same tests, all host threads */
        std::vector<std::vector<uint32_t>> part(nthreads);
        parallel_for(n1, nthreads, [&](unsigned tid, size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) {
                const lf_align_task &t = t1[i];
                const int32_t ed = r1[i].edit_distance, ql = (int32_t)t.q_len, tl = (int32_t)t.t_len;
                bool hit;
                if (t.mode == LF_MODE_SHW) hit = ql > kClipLen && (1 - ((float)ed / ql)) < kClipSim;      /* :1840 / :2175 */
                else hit = abs(ql - tl) >= kSplitLen && (1 - ((float)ed / ql)) < kSplitSim;               /* :1952 */
                if (hit) part[tid].push_back((uint32_t)i);
            }
        });
        for (auto &v : part) trig.insert(trig.end(), v.begin(), v.end());
    }
    R->stats.round1_tasks = n1;
    const double tm2 = now_ms();

    /* ---------------- round 2: triggers -> extensions ---------------- */
    std::vector<lf_extend_task> e2;
    std::vector<uint32_t> e2_src;   /* per extension: 2 * round-1 task + (1: the reversed one of a split pair) */
    std::vector<ClipInfo> clips;
    std::vector<SplitInfo> splits;
    std::vector<int32_t> gap_split(gpu_emit ? 0 : total_seeds, -1);   /* per (chain, seed i): index into splits, host emit only */
    std::vector<uint32_t> dirty_list, clean_list;   /* chains a trigger fired for (ascending), and the others */
    {   /* the few thousand hits, in task (= chain, then head / gaps / tail) order: the order of the serial scan; with
         * them the long heads / tails round 1 only measured (long_end): whatever the clip test says, round 3 has
         * something to do for them */
        std::vector<uint32_t> ev;
        {
            std::vector<uint32_t> long_tasks;
            for (size_t c = 0; c < n_chains; c++) {
                if (!hascand[c]) continue;
                const ChainPlan &p = plan[c];
                const lf_seed *sd = seeds + chains[c].seed_off;
                if (p.head_task >= 0 && (int32_t)sd[0].qPos > kClipLen) long_tasks.push_back((uint32_t)p.head_task);
                if (p.tail_task >= 0) {
                    const uint32_t n = chains[c].n_seeds;
                    const uint32_t readLen = (uint32_t)(reads->offsets[chains[c].read_id + 1] - reads->offsets[chains[c].read_id]);
                    if ((int32_t)readLen - (int32_t)(sd[n - 1].qPos + sd[n - 1].len) > kClipLen) long_tasks.push_back((uint32_t)p.tail_task);
                }
            }
            ev.resize(trig.size() + long_tasks.size());
            ev.resize((size_t)(std::set_union(trig.begin(), trig.end(), long_tasks.begin(), long_tasks.end(), ev.begin()) - ev.begin()));
        }
        size_t k = 0;
        while (k < ev.size()) {
            const uint32_t c = (uint32_t)(std::upper_bound(task_base.begin(), task_base.end(), (uint64_t)ev[k]) - task_base.begin() - 1);
            const lf_chain &ch = chains[c];
            const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
            ChainPlan &p = plan[c];
            dirty_list.push_back(c);
            p.split_lo = (int32_t)splits.size();
            const lf_seed *sd = seeds + ch.seed_off;
            uint32_t gi = 0;   /* gaps are visited in order: the chain's hits are ascending too */
            int64_t gtask = (int64_t)task_base[c] + (p.head_task >= 0 ? 1 : 0) - 1;   /* task of the last gap passed */
            for (; k < ev.size() && ev[k] < task_base[c + 1]; k++) {
                const int32_t ti = (int32_t)ev[k];
                const bool fired = std::binary_search(trig.begin(), trig.end(), (uint32_t)ti);
                const lf_align_task t = task_at(c, (size_t)ti);
                if (ti == p.head_task) {
                    p.head_clip = (int32_t)clips.size();
                    clips.push_back(ClipInfo{ fired ? (int32_t)e2.size() : -1, -1, 0, 0 });
                    if (fired) { e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand | LF_F_REVERSE_BOTH, true)); e2_src.push_back(2u * (uint32_t)ti); }
                } else if (ti == p.tail_task) {
                    p.tail_clip = (int32_t)clips.size();
                    clips.push_back(ClipInfo{ fired ? (int32_t)e2.size() : -1, -1, 0, 0 });
                    if (fired) { e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand, true)); e2_src.push_back(2u * (uint32_t)ti); }
                } else {
                    for (;; gi++) {   /* the gap whose task this is: gaps get their tasks in order */
                        const int32_t ql = (int32_t)(sd[gi + 1].qPos - (sd[gi].qPos + sd[gi].len)), tl = (int32_t)(sd[gi + 1].tPos - (sd[gi].tPos + sd[gi].len));
                        if (ql > 0 && tl > 0 && ++gtask == ti) break;
                    }
                    if (!gpu_emit) gap_split[gap_base[c] + gi] = (int32_t)splits.size();
                    SplitInfo si; memset(&si, 0, sizeof si);
                    si.gap_i = gi; si.task = ti;
                    gi++;
                    si.seed_idx = (uint32_t)(ch.seed_off + si.gap_i); si.ext_f = (int32_t)e2.size(); si.ext_r = si.ext_f + 1;
                    si.t_first = si.t_mid_f = si.t_mid_r = si.t_second = -1; si.split = false;
                    splits.push_back(si);
                    e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand, false));
                    e2.push_back(mk_ext(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, strand | LF_F_REVERSE_BOTH, false));
                    e2_src.push_back(2u * (uint32_t)ti); e2_src.push_back(2u * (uint32_t)ti + 1u);
                }
            }
            p.split_hi = (int32_t)splits.size();
        }
    }
    const double tq0 = now_ms();
    /* ---- chains no trigger fired for are final after round 1: their CIGAR / MD assembly (its own stream and
     *      host thread) overlaps rounds 2-3 ---- */
    EarlyEmit early;
    LfEmitDev E0;
    memset(&E0, 0, sizeof E0);
    size_t clean_slots = 0;
    if (gpu_emit) {
        DevState &d = ctx->devs[0];
        clean_list.reserve(n_chains - dirty_list.size());
        { size_t k = 0; for (size_t c = 0; c < n_chains; c++) { if (k < dirty_list.size() && dirty_list[k] == c) { k++; continue; } clean_list.push_back((uint32_t)c); clean_slots += chains[c].n_seeds + 1; } }
        E0.pac = d.pac.as<uint8_t>(); E0.chains = S.d_chains.as<lf_chain>(); E0.seeds = S.d_seeds.as<lf_seed>();
        E0.read_off = d.read_off.as<uint64_t>(); E0.task_base = S.d_task_base.as<uint64_t>(); E0.slot_base = S.d_slot_base.as<uint64_t>();
        E0.guards = S.d_guards.as<uint8_t>();
        E0.r1 = d.res.as<lf_align_result>(); E0.ops1 = d.ops.as<uint32_t>();   /* round 3 will run in the other pair of buffers */
        E0.slot_info = S.d_slot_info.as<LfSlotInfo>(); E0.slot_task = S.d_slot_task.as<uint32_t>();
        EmitSet &es = S.es[0];
#ifndef LF_EMU
        if (!es.have_stream) { if (cudaStreamCreateWithFlags(&es.st, cudaStreamNonBlocking) != cudaSuccess) { delete R; return LF_ERR_CUDA; } es.have_stream = true; }
#endif
        const size_t dirty_slots = n_slots - clean_slots;
        uint32_t *clean_pinned = (uint32_t *)stage_get(d, clean_list.size() * 4 + 16);
        if (!clean_pinned) { delete R; return LF_ERR_NOMEM; }
        if (!clean_list.empty()) memcpy(clean_pinned, clean_list.data(), clean_list.size() * 4);
        auto run_early = [&, dirty_slots, clean_pinned]() {
            if (set_dev(d)) { early.rc = LF_ERR_CUDA; return; }
            LfEmitDev E = E0;
            if ((early.rc = emit_size(d, es, E, clean_pinned, clean_list.size())) != 0) return;
            trace_mark(ctx, "early emit sized", es.st);
            /* one text buffer for both lists: the late list is appended, its size estimated from this one */
            const size_t bytes = es.ncig + es.nmd;
            const size_t est_late = dirty_slots ? (size_t)((double)bytes * ((double)dirty_slots / (double)(clean_slots ? clean_slots : 1)) * 1.5) + (1u << 16) : 0;
            if (io && bytes + est_late + 1 <= io->cap) { early.text = io->arena + io->off; early.cap = io->cap; early.borrowed = true; }
            else early.text = (char *)g_result_pool_pinned.get(bytes + est_late + 1, &early.cap);
            if (!early.text) { early.rc = LF_ERR_NOMEM; return; }
            if ((early.rc = emit_write(es, E, early.text, early.borrowed ? io->off : 0)) != 0) return;
            trace_mark(ctx, "early emit written", es.st);
            if (lfb_sync(es.st)) early.rc = LF_ERR_CUDA;
        };
#ifndef LF_EMU
        early.th = std::thread(run_early);
#else
        run_early();   /* the emulator is single-threaded */
#endif
    }
    const double tq1 = now_ms();
    std::vector<lf_extend_result> x2(e2.size());
    bool have_x2 = e2.empty();
    if (spec && !e2.empty()) {   /* the extensions have been running since before round 1: pick the results */
        LF_CH(spec_extend_wait(ctx));
        have_x2 = true;
        for (size_t k = 0; k < e2.size() && have_x2; k++) {
            const auto it = std::lower_bound(cand_ti.begin(), cand_ti.end(), e2_src[k] >> 1);
            if (it == cand_ti.end() || *it != (e2_src[k] >> 1)) { have_x2 = false; break; }   /* cannot happen: the triggers imply the length tests */
            x2[k] = sx2[cand_ext[(size_t)(it - cand_ti.begin())] + (e2_src[k] & 1u)];
        }
    } else if (spec) LF_CH(spec_extend_wait(ctx));
    if (!have_x2) {
        LF_CH(lf_gpu_upload_extend_tasks(ctx, e2.data(), e2.size()));
        LF_CH(lf_gpu_run_extend(ctx));
        LF_CH(lf_gpu_download_extend(ctx, x2.data()));
    }
    R->stats.round2_extends = e2.size();
    const double tq2 = now_ms();

    /* ---------------- round 3: follow-up alignments ---------------- */
    std::vector<lf_align_task> t3;
    for (const uint32_t c : dirty_list) {
        ChainPlan &p = plan[c];
        const lf_chain &ch = chains[c];
        const lf_seed *s = seeds + ch.seed_off;
        const uint32_t n = ch.n_seeds;
        const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
        if (p.head_clip >= 0) {
            ClipInfo &ci = clips[(size_t)p.head_clip];
            const lf_align_task t = task_at(c, (size_t)p.head_task);
            if (ci.ext >= 0) { ci.qle = x2[(size_t)ci.ext].qle; ci.tle = x2[(size_t)ci.ext].tle; }
            ci.t3 = (int32_t)t3.size();
            if (ci.ext >= 0 && ci.qle > 0 && ci.qle < (int32_t)t.q_len)                            /* :1850-1853 */
                t3.push_back(mk_task(ch.read_id, t.q_len - (uint32_t)ci.qle, (uint32_t)ci.qle, s[0].tPos - (uint32_t)ci.tle, (uint32_t)ci.tle,
                                     strand | LF_F_REVERSE_BOTH, LF_MODE_NW));
            else {   /* "use the already calculated results" (:1869-1886): the prefix-mode alignment itself, now with its path */
                ci.qle = (int32_t)t.q_len; ci.tle = (int32_t)t.t_len;
                t3.push_back(mk_task(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, t.flags & ~(unsigned)LF_F_NO_PATH, LF_MODE_SHW));
            }
        }
        for (int32_t sx = p.split_lo; sx < p.split_hi; sx++) {
            SplitInfo &si = splits[(size_t)sx];
            const lf_align_task t = task_at(c, (size_t)si.task);
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
            const lf_align_task t = task_at(c, (size_t)p.tail_task);
            if (ci.ext >= 0) { ci.qle = x2[(size_t)ci.ext].qle; ci.tle = x2[(size_t)ci.ext].tle; }
            ci.t3 = (int32_t)t3.size();
            if (ci.ext >= 0 && ci.qle > 0 && ci.qle < (int32_t)t.q_len)                            /* :2181-2184 */
                t3.push_back(mk_task(ch.read_id, t.q_off, (uint32_t)ci.qle, t.t_off, (uint32_t)ci.tle, strand, LF_MODE_NW));
            else {   /* :2196-2217 */
                ci.qle = (int32_t)t.q_len; ci.tle = (int32_t)t.t_len;
                t3.push_back(mk_task(ch.read_id, t.q_off, t.q_len, t.t_off, t.t_len, t.flags & ~(unsigned)LF_F_NO_PATH, LF_MODE_SHW));
            }
        }
    }
    const double tq3 = now_ms();
    const size_t n3 = t3.size();
    const size_t cap3 = lf_gpu_ops_capacity(t3.data(), n3);
    lf_align_result *r3 = (lf_align_result *)S.r3.reserve((n3 + 1) * sizeof(lf_align_result));
    uint8_t *ops3 = gpu_emit ? nullptr : (uint8_t *)S.ops3.reserve(cap3 + 64);
    if (!r3 || (!gpu_emit && !ops3)) { delete R; return LF_ERR_NOMEM; }
    /* keep the round-1 results and op stream in HBM: round 3 gets its own pair of buffers; whatever way the call ends,
     * the pairs are swapped back */
    struct SwapGuard { DevState *d; ~SwapGuard() { if (d) { std::swap(d->res, d->res_keep); std::swap(d->ops, d->ops_keep); } } } swap_guard{nullptr};
    if (gpu_emit) {
        DevState &d = ctx->devs[0];
        std::swap(d.res, d.res_keep); std::swap(d.ops, d.ops_keep);
        swap_guard.d = &d;
    }
    double t3a = tq3, t3b = tq3;
    if (n3) {
        if (gpu_emit) LF_CH(upload_align_tasks_k(ctx, t3.data(), n3));
        else LF_CH(lf_gpu_upload_align_tasks(ctx, t3.data(), n3));
        t3a = now_ms();
        LF_CH(lf_gpu_run_align(ctx));
        t3b = now_ms();
        for (DevState &dd : ctx->devs) if (dd.cls_count[LF_CLS_BAD]) { delete R; return fail(ctx, LF_ERR_BAD_ARG, "a follow-up alignment task is outside its read or the reference"); }
        trace_mark(ctx, "round-3 kernels done", ctx->devs[0].stream);
        if (!gpu_emit) LF_CH(lf_gpu_download_align(ctx, r3, ops3, cap3));
        else { DevState &d = ctx->devs[0]; if (lfb_d2h(r3, d.res.p, n3 * sizeof(lf_align_result), d.stream)) { delete R; return LF_ERR_CUDA; } }
    }
    LF_CH(lf_gpu_sync(ctx));
    R->stats.round3_tasks = n3;
    const double tm3 = now_ms();
    if (getenv("LF_CHAIN_TRACE")) fprintf(stderr, "[lf_chain %p +%.2f] ", (void *)ctx, now_ms() - g_trace_t0), fprintf(stderr, "r23: trigger scan %.2f, early-emit start %.2f, round 2 %.2f, round-3 tasks %.2f, round 3 %.2f (upload %.2f, prep + sort + class counts + launches %.2f, kernels + results %.2f)\n", tq0 - tm2, tq1 - tq0, tq2 - tq1, tq3 - tq2, tm3 - tq3, t3a - tq3, t3b - t3a, tm3 - t3b);
#undef LF_CH

    if (gpu_emit) {
        /* ---------------- late emit on the GPU: the chains rounds 2-3 worked on ---------------- */
        DevState &d = ctx->devs[0];
        lfb_stream st = d.stream;
        EmitSet &es0 = S.es[0], &es1 = S.es[1];
        es1.st = st;
#define LF_G(expr) do { if ((expr) != 0) { delete R; return fail(ctx, LF_ERR_CUDA, #expr); } } while (0)
        LfEmitDev E1 = E0;
        const size_t n_dirty = dirty_list.size();
        if (n_dirty) {
            std::vector<int32_t> clipv(4 * n_chains, -1);
            std::vector<uint32_t> split_begin(n_chains + 1, 0);
            std::vector<LfSplitDev> sdev;
            size_t k = 0;
            for (size_t c = 0; c < n_chains; c++) {
                split_begin[c] = (uint32_t)sdev.size();
                if (k >= n_dirty || dirty_list[k] != c) continue;
                k++;
                const ChainPlan &p = plan[c];
                if (p.head_clip >= 0) { const ClipInfo &ci = clips[(size_t)p.head_clip]; clipv[4 * c + 0] = ci.t3; clipv[4 * c + 1] = ci.qle; }
                if (p.tail_clip >= 0) { const ClipInfo &ci = clips[(size_t)p.tail_clip]; clipv[4 * c + 2] = ci.t3; clipv[4 * c + 3] = ci.qle; }
                for (int32_t sx = p.split_lo; sx < p.split_hi; sx++) {
                    const SplitInfo &si = splits[(size_t)sx];
                    LfSplitDev v; memset(&v, 0, sizeof v);
                    v.gap_i = si.gap_i; v.t_first = si.t_first; v.t_mid_r = si.t_mid_r; v.t_second = si.t_second;
                    v.qs2 = si.qs2; v.ts2 = si.ts2; v.qe2 = si.qe2; v.te2 = si.te2; v.split = si.split ? 1u : 0u; v.inv = 0;
                    if (si.split && si.t_mid_f >= 0) {
                        const lf_align_result &rf = r3[(size_t)si.t_mid_f], &rr = r3[(size_t)si.t_mid_r];
                        const int32_t ql2 = (int32_t)(si.qe2 - si.qs2);
                        if ((1 - ((double)rr.edit_distance / ql2)) > (1 - ((double)rf.edit_distance / ql2))
                            && (1 - ((double)rr.edit_distance / ql2)) > kReverseSim) v.inv = 1;                /* :2040-2041 */
                    }
                    sdev.push_back(v);
                }
            }
            split_begin[n_chains] = (uint32_t)sdev.size();
            LF_G(S.d_clip.reserve(4 * n_chains * 4 + 64)); LF_G(S.d_split_begin.reserve((n_chains + 1) * 4 + 64)); LF_G(S.d_splits.reserve((sdev.size() + 1) * sizeof(LfSplitDev) + 64));
            LF_G(h2d_k(d, S.d_clip.p, clipv.data(), 4 * n_chains * 4, st, false)); LF_G(h2d_k(d, S.d_split_begin.p, split_begin.data(), (n_chains + 1) * 4, st, false));
            if (!sdev.empty()) LF_G(h2d_k(d, S.d_splits.p, sdev.data(), sdev.size() * sizeof(LfSplitDev), st, false));
            E1.clip = S.d_clip.as<int32_t>(); E1.split_begin = S.d_split_begin.as<uint32_t>(); E1.splits = S.d_splits.as<LfSplitDev>();
            E1.r3 = d.res.as<lf_align_result>(); E1.ops3 = d.ops.as<uint32_t>();
            uint32_t *dirty_pinned = (uint32_t *)stage_get(d, n_dirty * 4 + 16);
            if (!dirty_pinned) { delete R; return LF_ERR_NOMEM; }
            memcpy(dirty_pinned, dirty_list.data(), n_dirty * 4);
            LF_G(emit_size(d, es1, E1, dirty_pinned, n_dirty));
        } else { es1.n = es1.nrec = es1.ncig = es1.nmd = 0; }
        const double te1 = now_ms();
        early.join();
        const double te2 = now_ms();
        if (early.rc != 0) { delete R; return fail(ctx, early.rc, "early emit"); }
        const size_t bytes0 = es0.n ? es0.ncig + es0.nmd : 0, bytes1 = es1.ncig + es1.nmd;
        R->pinned = true;
        R->text = early.text; R->text_cap = early.cap; R->text_borrowed = early.borrowed; early.text = nullptr;
        if (!R->text && io && bytes1 + 1 <= io->cap) { R->text = io->arena + io->off; R->text_cap = io->cap; R->text_borrowed = true; }   /* no early list */
        if (!R->text || bytes0 + bytes1 + 1 > R->text_cap) {   /* the estimate fell short (or there was no early list): move to a bigger buffer */
            size_t cap = 0;
            char *nt = (char *)g_result_pool_pinned.get(bytes0 + bytes1 + 1, &cap);
            if (!nt) { delete R; return LF_ERR_NOMEM; }
            if (R->text) {
                parallel_for(bytes0, nthreads, [&](unsigned, size_t lo, size_t hi) { memcpy(nt + lo, R->text + lo, hi - lo); }, 1 << 20);
                if (!R->text_borrowed) g_result_pool_pinned.put(R->text, R->text_cap);
                else {   /* the early records point into the arena: re-base them to the lane's own buffer */
                    lf_sam_record *a = (lf_sam_record *)es0.h_recs.p;
                    for (size_t k = 0; k < (es0.n ? es0.nrec : 0); k++) { a[k].cigar_off -= io->off; a[k].md_off -= io->off; }
                }
            }
            R->text = nt; R->text_cap = cap; R->text_borrowed = false;
        }
        if (n_dirty) { LF_G(emit_write(es1, E1, R->text + bytes0, (R->text_borrowed ? io->off : 0) + bytes0)); trace_mark(ctx, "late emit written", st); LF_G(lfb_sync(st)); }
        const double te3 = now_ms();
        if (getenv("LF_CHAIN_TRACE")) fprintf(stderr, "[lf_chain %p +%.2f] late emit: inputs + size pass %.2f, wait for the early emit %.2f, write pass + download %.2f (%zu bytes)\n", (void *)ctx, now_ms() - g_trace_t0, te1 - tm3, te2 - te1, te3 - te2, bytes1);
        LF_G(lfb_last_error());
#undef LF_G
        /* records of the two lists, merged back into chain order */
        const size_t nrec = (es0.n ? es0.nrec : 0) + es1.nrec;
        R->n_recs = nrec; R->text_bytes = bytes0 + bytes1;
        R->recs = (lf_sam_record *)g_result_pool_pinned.get((nrec + 1) * sizeof(lf_sam_record), &R->recs_cap);
        if (!R->recs) { delete R; return LF_ERR_NOMEM; }
        {
            const lf_sam_record *a = (const lf_sam_record *)es0.h_recs.p, *b = (const lf_sam_record *)es1.h_recs.p;
            const size_t na = es0.n ? es0.nrec : 0, nb = es1.nrec;
            size_t ia = 0, ib = 0, o = 0;
            while (ia < na && ib < nb) {
                if (a[ia].chain_id < b[ib].chain_id) { const uint32_t c = a[ia].chain_id; while (ia < na && a[ia].chain_id == c) R->recs[o++] = a[ia++]; }
                else { const uint32_t c = b[ib].chain_id; while (ib < nb && b[ib].chain_id == c) R->recs[o++] = b[ib++]; }
            }
            if (ia < na) { memcpy(R->recs + o, a + ia, (na - ia) * sizeof(lf_sam_record)); o += na - ia; }
            if (ib < nb) { memcpy(R->recs + o, b + ib, (nb - ib) * sizeof(lf_sam_record)); o += nb - ib; }
        }
        /* (swap_guard hands the bigger pair back to the next round 1) */
        ctx->stats.kernel_launches = lfb_launches;
        R->stats.records = nrec;
        const double tm4 = now_ms();
        R->stats.ms_tasks = (float)(tm1 - tm0); R->stats.ms_round1 = (float)(tm2 - tm1); R->stats.ms_rounds23 = (float)(tm3 - tm2); R->stats.ms_emit = (float)(tm4 - tm3); R->stats.ms_merge = 0.f;
        if (getenv("LF_CHAIN_TRACE")) fprintf(stderr, "[lf_chain %p +%.2f] ", (void *)ctx, now_ms() - g_trace_t0), fprintf(stderr, "gpu emit: tasks %.2f r1 %.2f r23 %.2f late emit %.2f (chains %zu early + %zu late, recs %zu, text %zu MB)\n",
                                              tm1 - tm0, tm2 - tm1, tm3 - tm2, tm4 - tm3, clean_list.size(), n_dirty, nrec, (bytes0 + bytes1) >> 20);
        if (!io) trace_print();
        *out = R;
        return LF_OK;
    }

    /* ---------------- emit on host threads: the reference's accumulation, chain by chain ---------------- */
    if (n_chains < 256) nthreads = 1; /* same threshold as parallel_for's serial mode: one part, one slice */
    if (!S.parts) S.parts = new std::vector<Emit>();
    std::vector<Emit> &parts = *S.parts;   /* kept between calls: their buffers are already mapped */
    if (parts.size() != nthreads) parts.resize(nthreads);
    /* Upper bound of a chain's text: every non-match op and every run boundary costs at most 16 bytes in
     * CIGAR + MD together; non-match ops = the edit distances, run boundaries <= 2 per piece.  The result
     * buffer is sliced by these bounds so that the emit threads write their records in place (no merge). */
    std::vector<uint64_t> tbound(n_chains + 1, 0);
    for (size_t c = 0; c < n_chains; c++) {
        uint64_t ev = 4ull * chains[c].n_seeds + 64;
        if (plan[c].head_task >= 0) ev += (uint64_t)r1[(size_t)plan[c].head_task].edit_distance;
        if (plan[c].tail_task >= 0) ev += (uint64_t)r1[(size_t)plan[c].tail_task].edit_distance;
        for (uint64_t g = gap_base[c]; g < gap_base[c + 1]; g++) {
            if (gap_task[g] >= 0) ev += (uint64_t)r1[(size_t)gap_task[g]].edit_distance;
            if (gap_split[g] >= 0) { /* a split re-aligns its pieces: distances of the round-3 tasks, plus up to three records */
                const SplitInfo &si = splits[(size_t)gap_split[g]];
                const int32_t ids[4] = { si.t_first, si.t_mid_f, si.t_mid_r, si.t_second };
                for (int32_t id : ids) if (id >= 0) ev += (uint64_t)r3[(size_t)id].edit_distance + r3[(size_t)id].ops_len / 4 + 64;
                ev += 256;
            }
        }
        const lf_seed *sd = seeds + chains[c].seed_off;   /* pure insert / delete gaps print one item per deleted base */
        ev += (uint64_t)(sd[chains[c].n_seeds - 1].tPos - sd[0].tPos) / 8;
        if (plan[c].head_clip >= 0 && clips[(size_t)plan[c].head_clip].t3 >= 0) ev += (uint64_t)r3[(size_t)clips[(size_t)plan[c].head_clip].t3].edit_distance;
        if (plan[c].tail_clip >= 0 && clips[(size_t)plan[c].tail_clip].t3 >= 0) ev += (uint64_t)r3[(size_t)clips[(size_t)plan[c].tail_clip].t3].edit_distance;
        tbound[c + 1] = tbound[c] + 16 * ev;
    }
    R->text_bytes = tbound[n_chains];
    R->text = (char *)g_result_pool.get(R->text_bytes + 1, &R->text_cap);
    if (!R->text) { delete R; return LF_ERR_NOMEM; }
    for (unsigned t = 0; t < nthreads; t++) {
        Emit &E = parts[t];
        E.recs.clear(); E.overflow = false;
        E.base = R->text;
        E.cur = R->text + tbound[n_chains * t / nthreads];
        E.lim = R->text + tbound[n_chains * (t + 1) / nthreads];
    }
    auto work = [&](unsigned tid, size_t c_lo, size_t c_hi) {
        Emit &E = parts[tid];
        RecBuf &B = E.B;
        for (size_t c = c_lo; c < c_hi; c++) {
            const lf_chain &ch = chains[c];
            const lf_seed *s = seeds + ch.seed_off;
            const uint32_t n = ch.n_seeds;
            const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
            const ChainPlan &p = plan[c];
            const uint32_t flag_norm = ch.is_rev ? 16u : 0u, flag_opp = ch.is_rev ? 0u : 16u;
            uint32_t flag = flag_norm, pos = s[0].tPos, qStart = s[0].qPos, posEnd = 0, qEnd = 0;
            int32_t editScore = 0;
            /* every op of a record consumes a read base or a reference base of the chain's span (+ clips) */
            B.begin(3 * (size_t)readLen + (size_t)(s[n - 1].tPos + s[n - 1].len - s[0].tPos) + 4096);
            /* head (:1820-1899) */
            const int32_t a = (int32_t)s[0].qPos;
            if (a > 0) {
                if (p.head_guard) {
                    const ClipInfo *ci = p.head_clip >= 0 ? &clips[(size_t)p.head_clip] : nullptr;
                    if (ci && ci->t3 >= 0) {
                        const lf_align_result &r = r3[(size_t)ci->t3];
                        B.run('I', (size_t)(a - ci->qle));
                        B.segment(ops3, r, true, pac_host, s[0].tPos - (uint32_t)(r.end_location + 1));
                        editScore -= r.edit_distance;
                        pos = s[0].tPos - (uint32_t)r.end_location - 1;
                        qStart = s[0].qPos - (uint32_t)ci->qle;
                    } else {
                        const lf_align_result &r = r1[(size_t)p.head_task];
                        B.segment(ops1, r, true, pac_host, s[0].tPos - (uint32_t)(r.end_location + 1));
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
                            B.segment(ops3, r, false, pac_host, ts);
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
                                SlowRec Q;
                                Q.run('I', (size_t)si->qs2);
                                Q.segment(ops3, rr, pac_host, si->ts2);
                                Q.cig.append((size_t)(readLen - si->qe2), 'I');
                                Q.md.insert((size_t)0, (size_t)(readLen - si->qe2), '-');              /* sic, :2056-2057 */
                                E.tc.clear(); E.tm.clear();
                                Q.finish(E.tc, E.tm);
                                E.push_strings((uint32_t)c, flag_opp, si->ts2, si->te2, si->qs2, si->qe2, -rr.edit_distance);
                            }
                        }
                        B.run('I', (size_t)si->qe2);
                        if (si->t_second >= 0) {
                            const lf_align_result &r = r3[(size_t)si->t_second];
                            B.segment(ops3, r, true, pac_host, si->te2);
                            editScore -= r.edit_distance;
                        }
                        flag = flag_norm; pos = si->te2; qStart = si->qe2;
                        numAnchorsSoFar = 0;
                    } else {
                        const lf_align_result &r = r1[(size_t)gt];
                        editScore -= r.edit_distance;
                        B.segment(ops1, r, false, pac_host, ts);
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
                        B.segment(ops3, r, false, pac_host, ts);
                        editScore -= r.edit_distance;
                        posEnd = ts + (uint32_t)r.end_location;
                        qEnd = qs + (uint32_t)ci->qle;
                        B.run('I', (size_t)(b - ci->qle));
                    } else {
                        const lf_align_result &r = r1[(size_t)p.tail_task];
                        editScore -= r.edit_distance;
                        B.segment(ops1, r, false, pac_host, ts);
                        posEnd = ts + (uint32_t)r.end_location;
                        qEnd = readLen;
                    }
                } else B.run('I', (size_t)b);
            }
            E.push((uint32_t)c, flag, pos, posEnd, qStart, qEnd, editScore, B);
        }
    };
    parallel_for(n_chains, nthreads, work, 256);
    const double tm3b = now_ms();
    /* records of all threads, in chain order (the text is already in place) */
    size_t nrec = 0;
    bool overflow = false;
    for (Emit &E : parts) { nrec += E.recs.size(); overflow |= E.overflow; }
    if (overflow) { delete R; return fail(ctx, LF_ERR_NOMEM, "chain text bound exceeded"); }
    R->n_recs = nrec;
    R->recs = (lf_sam_record *)g_result_pool.get((nrec + 1) * sizeof(lf_sam_record), &R->recs_cap);
    if (!R->recs) { delete R; return LF_ERR_NOMEM; }
    nrec = 0;
    for (Emit &E : parts) { if (!E.recs.empty()) memcpy(R->recs + nrec, E.recs.data(), E.recs.size() * sizeof(lf_sam_record)); nrec += E.recs.size(); }
    R->stats.records = R->n_recs;
    const double tm4 = now_ms();
    if (getenv("LF_CHAIN_TRACE")) fprintf(stderr, "[lf_chain %p +%.2f] ", (void *)ctx, now_ms() - g_trace_t0), fprintf(stderr, "tasks %.2f r1 %.2f r23 %.2f emit %.2f gather %.2f (threads %u, recs %zu, text bound %zu MB)\n",
                                          tm1 - tm0, tm2 - tm1, tm3 - tm2, tm3b - tm3, tm4 - tm3b, nthreads, R->n_recs, R->text_bytes >> 20);
    R->stats.ms_tasks = (float)(tm1 - tm0); R->stats.ms_round1 = (float)(tm2 - tm1); R->stats.ms_rounds23 = (float)(tm3 - tm2); R->stats.ms_emit = (float)(tm3b - tm3); R->stats.ms_merge = (float)(tm4 - tm3b);
    *out = R;
    return LF_OK;
}

/* One lane per sub-batch: a single-device child context with its own streams, buffers and host thread.  Lane j lives on
 * device j % n_devices and borrows that device's copy of the reference. */
static lf_gpu_ctx *lane_ctx(lf_gpu_ctx *ctx, size_t j)
{
    while (ctx->lanes.size() <= j) {
        DevState &pd = ctx->devs[ctx->lanes.size() % ctx->devs.size()];
        lf_gpu_ctx *c = new lf_gpu_ctx();
        c->rt = ctx->rt;
        c->l_pac = ctx->l_pac;
        c->devs.resize(1);
        DevState &d = c->devs[0];
        d.dev = pd.dev;
        d.pac.p = pd.pac.p; d.pac.cap = pd.pac.cap; d.pac_borrowed = true;
        if (set_dev(d) || init_dev_streams(d) || !(d.pinned = lfb_host_alloc(sizeof(HostTotals)))) { lf_gpu_destroy(c); return nullptr; }
        ctx->lanes.push_back(c);
    }
    return ctx->lanes[j];
}

/* The call the reference side makes.  A chunk is cut into sub-batches that run as a software pipeline, each on a lane:
 *
 *   the slow lane   the chains that can keep a lane busy for milliseconds whatever the batch size -- a gap, head or tail
 *                   long enough for the lane-group / large kernels or for a clip / split trigger (src/LordFAST.cpp:1840,
 *                   :1952, :2175), i.e. every chain that can reach rounds 2 and 3.  A few percent of a chunk; their reads
 *                   are gathered into pinned staging and go to the device FIRST, so that their long chain of dependent
 *                   launches (round 1 with 1 - 2 kbp tasks, extensions, round 3, the walk over long garbage alignments in
 *                   the emit) runs while the bulk of the reads is still crossing PCIe;
 *   the fast lanes  the other chains, as consecutive ranges balanced by seeds; no task above 512 rows (14 % of a config-2
 *                   chunk's chains hold one; at 256 rows it would be 86 %), no trigger
 *                   candidate, so a lane is upload -> round 1 -> emit.  Their reads follow in lane order on one stream:
 *                   the first lane aligns while the others' reads are in flight, and the CIGAR / MD text of a lane goes
 *                   back over PCIe (k_emit_slots writes it straight to pinned memory) while the next lanes compute.
 *
 * With several devices in the context the lanes are spread over them: reads sharded by contiguous ranges, every device
 * holding the whole reference, records merged in chain order -- the data-parallel loop of src/LordFAST.cpp:295-316.
 * LF_CHAIN_LANES overrides the number of lanes (1: no pipeline), LF_CHAIN_SLOW_LANE=0 switches the slow lane off. */
int lf_gpu_align_chains(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_contigs *contigs, const lf_seed *seeds,
                        const lf_chain *chains, size_t n_chains, const uint8_t *pac_host, lf_chain_results **out)
{
    if (!ctx || !reads || !reads->offsets || !contigs || !seeds || (!chains && n_chains) || !pac_host || !out || contigs->n < 1) return LF_ERR_BAD_ARG;
    const size_t ndev = ctx->devs.size();
    /* One device: one lane by default.  Measured on a config-2 chunk (profiles/r03_e2e_lanes.txt): the dependent launches of a
     * lane (round 1 with its 1.5 kbp tasks, extensions, round 3, emit) take ~9 ms whatever the batch size, so 2 - 8 lanes
     * only moved a call from 15.2 to 14.0 - 14.9 ms; callers that want the PCIe link busy keep two calls in flight on two
     * contexts instead (10.5 ms per chunk).  Several devices: a lane per device. */
    size_t K = ndev;
    if (const char *e = getenv("LF_CHAIN_LANES")) { if (atoi(e) > 0) K = (size_t)atoi(e); }
    if (K > 64) K = 64;
    if (K > n_chains) K = n_chains ? n_chains : 1;
    if (K < 1) K = 1;
    g_trace_t0 = now_ms();
    if (ndev == 1 && K == 1) return align_chains_one(ctx, reads, contigs, seeds, chains, n_chains, pac_host, out, nullptr);
    if (!reads->bases) return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_align_chains: resident reads (bases == NULL) need a single device and a single lane");
    *out = nullptr;
    for (size_t c = 0; c < n_chains; c++) if (chains[c].read_id >= reads->n_reads || chains[c].n_seeds < 2) return LF_ERR_BAD_ARG;
    ChainScratch &S = chain_scratch(ctx);
    if (!S.workers) S.workers = new LfWorkers();
    struct WorkersScope { LfWorkers *prev; WorkersScope(LfWorkers *w) : prev(tl_workers) { tl_workers = w; } ~WorkersScope() { tl_workers = prev; } } workers_scope(S.workers);
    unsigned nthreads = std::thread::hardware_concurrency();
    {
        const char *e = getenv("LF_HOST_THREADS"), *lw = getenv("LOCAL_WORLD_SIZE");
        if (e && atoi(e) > 0) nthreads = (unsigned)atoi(e);
        else if (lw && atoi(lw) > 1) nthreads = nthreads / (unsigned)atoi(lw);
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;

    /* ---- which chains go to the slow lane ---- */
    std::vector<uint8_t> slow(n_chains, 0);
    const bool want_slow = K >= 2 && !(getenv("LF_CHAIN_SLOW_LANE") && atoi(getenv("LF_CHAIN_SLOW_LANE")) == 0);
    size_t n_slow = 0;
    if (want_slow) {
        parallel_for(n_chains, nthreads, [&](unsigned, size_t lo, size_t hi) {
            for (size_t c = lo; c < hi; c++) {
                const lf_chain &ch = chains[c];
                const lf_seed *s = seeds + ch.seed_off;
                const uint32_t n = ch.n_seeds;
                const uint32_t readLen = (uint32_t)(reads->offsets[ch.read_id + 1] - reads->offsets[ch.read_id]);
                bool sl = s[0].qPos > 512u || (int64_t)readLen - (int64_t)(s[n - 1].qPos + s[n - 1].len) > 512;
                for (uint32_t i = 0; i + 1 < n && !sl; i++) {
                    const int64_t ql = (int64_t)s[i + 1].qPos - (int64_t)(s[i].qPos + s[i].len), tl = (int64_t)s[i + 1].tPos - (int64_t)(s[i].tPos + s[i].len);
                    sl = ql > 512 || tl > 768 || (ql > 0 && tl > 0 && (ql - tl >= kSplitLen || tl - ql >= kSplitLen));
                }
                slow[c] = sl ? 1 : 0;
            }
        }, 1024);
        for (size_t c = 0; c < n_chains; c++) n_slow += slow[c];
        if (n_slow * 3 > n_chains || n_slow == 0) { n_slow = 0; std::fill(slow.begin(), slow.end(), 0); }   /* mostly slow chains: plain consecutive lanes */
    }
    const bool have_slow = n_slow > 0;
    const size_t KF = have_slow ? K - 1 : K;      /* fast lanes */

    struct Sub {
        std::vector<uint32_t> ids;       /* global index of the lane's chains, ascending */
        std::vector<lf_chain> ch;        /* ... re-based to the lane's reads and seeds */
        std::vector<lf_seed> own_seeds;  /* slow lane: its chains' seeds, gathered */
        const lf_seed *seeds = nullptr;
        lf_reads rd; size_t est = (size_t)1 << 18;
        const uint8_t *src = nullptr; uint64_t nbytes = 0; uint64_t *offs = nullptr;
        lf_chain_results *R = nullptr; int rc = 0; LaneIO io;
    };
    std::vector<Sub> sub(K);
    /* fast lanes: consecutive chains, balanced by seeds */
    {
        uint64_t total = 0;
        for (size_t c = 0; c < n_chains; c++) if (!slow[c]) total += chains[c].n_seeds;
        size_t c = 0; uint64_t acc = 0;
        for (size_t k = 0; k < KF; k++) {
            Sub &s = sub[(have_slow ? 1 : 0) + k];
            const uint64_t target = total * (k + 1) / KF;
            while (c < n_chains && (acc < target || k + 1 == KF)) { if (!slow[c]) { s.ids.push_back((uint32_t)c); acc += chains[c].n_seeds; } c++; }
        }
        if (have_slow) for (size_t cc = 0; cc < n_chains; cc++) if (slow[cc]) sub[0].ids.push_back((uint32_t)cc);
    }
    {   /* drop empty lanes */
        std::vector<Sub> keep;
        for (Sub &s : sub) if (!s.ids.empty()) keep.push_back(std::move(s));
        sub.swap(keep);
        K = sub.size();
    }
    /* ---- per lane: reads it touches, chains re-based, upload source ---- */
    size_t noff = 0;
    uint64_t lane_bytes = 0, slow_bytes = 0;
    std::vector<uint32_t> slow_rid;                 /* slow lane: the distinct reads of its chains, ascending */
    for (size_t k = 0; k < K; k++) {
        Sub &s = sub[k];
        const bool is_slow = have_slow && k == 0;
        if (is_slow) {
            std::vector<uint32_t> r;
            r.reserve(s.ids.size());
            for (uint32_t c : s.ids) r.push_back(chains[c].read_id);
            std::sort(r.begin(), r.end());
            r.erase(std::unique(r.begin(), r.end()), r.end());
            slow_rid.swap(r);
            for (uint32_t rid : slow_rid) slow_bytes += reads->offsets[rid + 1] - reads->offsets[rid];
            s.nbytes = slow_bytes; s.rd.n_reads = (uint32_t)slow_rid.size();
            noff += slow_rid.size() + 1;
        } else {
            uint32_t rmin = 0xffffffffu, rmax = 0;
            for (uint32_t c : s.ids) { const uint32_t r = chains[c].read_id; if (r < rmin) rmin = r; if (r > rmax) rmax = r; }
            s.src = reads->bases + reads->offsets[rmin];
            s.nbytes = reads->offsets[rmax + 1] - reads->offsets[rmin];
            s.rd.n_reads = rmax - rmin + 1;
            s.rd.bases = s.src;
            s.est = rmin;   /* parked: first read of the lane */
            noff += (size_t)s.rd.n_reads + 1;
            lane_bytes += s.nbytes;
        }
    }
    if (ndev == 1 && lane_bytes > reads->offsets[reads->n_reads] + reads->offsets[reads->n_reads] / 3)   /* chains not in read order: every lane would upload most of the reads */
        return align_chains_one(ctx, reads, contigs, seeds, chains, n_chains, pac_host, out, nullptr);
    uint64_t *offs = (uint64_t *)S.meta_stage.reserve(noff * 8 + 64);
    uint8_t *slow_stage = have_slow ? (uint8_t *)S.t1.reserve((size_t)slow_bytes + 64) : nullptr;
    if (!offs || (have_slow && !slow_stage)) return LF_ERR_NOMEM;
    {
        size_t o = 0;
        for (size_t k = 0; k < K; k++) {
            Sub &s = sub[k];
            const bool is_slow = have_slow && k == 0;
            s.offs = offs + o;
            s.ch.resize(s.ids.size());
            if (is_slow) {
                uint64_t b = 0;
                for (size_t i = 0; i < slow_rid.size(); i++) { s.offs[i] = b; b += reads->offsets[slow_rid[i] + 1] - reads->offsets[slow_rid[i]]; }
                s.offs[slow_rid.size()] = b;
                parallel_for(slow_rid.size(), nthreads, [&](unsigned, size_t lo, size_t hi) {
                    for (size_t i = lo; i < hi; i++) memcpy(slow_stage + s.offs[i], reads->bases + reads->offsets[slow_rid[i]], (size_t)(s.offs[i + 1] - s.offs[i]));
                }, 64);
                s.src = slow_stage; s.rd.bases = slow_stage;
                size_t ns = 0;
                for (uint32_t c : s.ids) ns += chains[c].n_seeds;
                s.own_seeds.resize(ns);
                size_t so = 0;
                s.est = (size_t)1 << 18;
                for (size_t i = 0; i < s.ids.size(); i++) {
                    const lf_chain &g = chains[s.ids[i]];
                    memcpy(s.own_seeds.data() + so, seeds + g.seed_off, (size_t)g.n_seeds * sizeof(lf_seed));
                    lf_chain &ch = s.ch[i];
                    ch = g; ch.seed_off = so;
                    ch.read_id = (uint32_t)(std::lower_bound(slow_rid.begin(), slow_rid.end(), g.read_id) - slow_rid.begin());
                    so += g.n_seeds;
                    s.est += (size_t)(reads->offsets[g.read_id + 1] - reads->offsets[g.read_id]);
                }
                s.seeds = s.own_seeds.data();
                o += slow_rid.size() + 1;
            } else {
                const uint32_t rmin = (uint32_t)s.est;
                const uint64_t b0 = reads->offsets[rmin];
                for (uint32_t r = 0; r <= s.rd.n_reads; r++) s.offs[r] = reads->offsets[rmin + r] - b0;
                uint64_t smin = (uint64_t)-1;
                for (uint32_t c : s.ids) if (chains[c].seed_off < smin) smin = chains[c].seed_off;
                s.est = (size_t)1 << 18;
                for (size_t i = 0; i < s.ids.size(); i++) {
                    const lf_chain &g = chains[s.ids[i]];
                    lf_chain &ch = s.ch[i];
                    ch = g; ch.seed_off = g.seed_off - smin; ch.read_id = g.read_id - rmin;
                    s.est += (size_t)(reads->offsets[g.read_id + 1] - reads->offsets[g.read_id]);   /* ~0.55 B of CIGAR + MD per aligned read base at 13 % error */
                }
                s.seeds = seeds + smin;
                o += (size_t)s.rd.n_reads + 1;
            }
            s.rd.offsets = s.offs;
        }
    }
    /* ---- one text buffer for the whole call, a slice per lane ---- */
    size_t arena_bytes = 0;
    for (Sub &s : sub) { s.io.off = arena_bytes; s.io.cap = s.est; arena_bytes += (s.est + 255) & ~(size_t)255; }
    size_t arena_cap = 0;
    char *arena = (char *)g_result_pool_pinned.get(arena_bytes + 1, &arena_cap);
    if (!arena) return LF_ERR_NOMEM;
    trace_ref(ctx->devs[0].stream);
    /* ---- the reads of every lane, in lane order (slow lane first), on the parent's stream of the lane's device ---- */
    int rc = LF_OK;
    for (size_t k = 0; k < K && rc == LF_OK; k++) {
        Sub &s = sub[k];
        lf_gpu_ctx *lc = lane_ctx(ctx, k);
        if (!lc) { rc = LF_ERR_NOMEM; break; }
        DevState &pd = ctx->devs[k % ndev], &d = lc->devs[0];
        const uint32_t nr = s.rd.n_reads;
        s.io.arena = arena; s.io.nthreads = nthreads / (unsigned)K ? nthreads / (unsigned)K : 1u;
        if (set_dev(d) || d.bases.reserve((size_t)s.nbytes + 64) || d.read_off.reserve(((size_t)nr + 1) * 8 + 64)
            || h2d_k(pd, d.read_off.p, s.offs, ((size_t)nr + 1) * 8, pd.stream, true) || lfb_h2d(d.bases.p, s.src, (size_t)s.nbytes, pd.stream)) { rc = LF_ERR_CUDA; break; }
#ifndef LF_EMU
        cudaEventRecord(d.up_ev, pd.stream);
#endif
        d.reads_preloaded = true;
    }
    /* ---- the lanes, a host thread each ---- */
    if (rc == LF_OK) {
        auto run_lane = [&](size_t k) {
            Sub &s = sub[k];
            s.rc = align_chains_one(ctx->lanes[k], &s.rd, contigs, s.seeds, s.ch.data(), s.ch.size(), pac_host, &s.R, &s.io);
        };
#ifndef LF_EMU
        std::vector<std::thread> th;
        for (size_t k = 1; k < K; k++) th.emplace_back(run_lane, k);
        run_lane(0);
        for (auto &t : th) t.join();
#else
        for (size_t k = 0; k < K; k++) run_lane(k);   /* the emulator is single-threaded */
#endif
        for (size_t k = 0; k < K; k++) if (sub[k].rc != LF_OK && rc == LF_OK) { rc = sub[k].rc; ctx->err = "lane " + std::to_string(k) + ": " + ctx->lanes[k]->err; }
    }
    for (size_t k = 0; k < ctx->lanes.size(); k++) ctx->lanes[k]->devs[0].reads_preloaded = false;
    trace_print();
    if (rc != LF_OK) {
        for (Sub &s : sub) delete s.R;
        g_result_pool_pinned.put(arena, arena_cap);
        return rc;
    }
    /* ---- merge: records in chain order, text already in place ---- */
    lf_chain_results *R = new lf_chain_results();
    memset(&R->stats, 0, sizeof R->stats);
    R->pinned = true;
    size_t nrec = 0, text_end = 0;
    bool all_in_arena = true;
    for (Sub &s : sub) { nrec += s.R->n_recs; all_in_arena &= s.R->text_borrowed || s.R->text_bytes == 0; }
    R->recs = (lf_sam_record *)g_result_pool_pinned.get((nrec + 1) * sizeof(lf_sam_record), &R->recs_cap);
    if (!R->recs) { for (Sub &s : sub) delete s.R; g_result_pool_pinned.put(arena, arena_cap); delete R; return LF_ERR_NOMEM; }
    if (!all_in_arena) {   /* a lane's text outgrew its slice (it moved to a buffer of its own): gather everything into a new buffer */
        size_t total = 0;
        for (Sub &s : sub) total += s.R->text_bytes;
        size_t cap = 0;
        char *nt = (char *)g_result_pool_pinned.get(total + 1, &cap);
        if (!nt) { for (Sub &s : sub) delete s.R; g_result_pool_pinned.put(arena, arena_cap); delete R; return LF_ERR_NOMEM; }
        size_t o = 0;
        for (Sub &s : sub) {
            const char *src = s.R->text_borrowed ? arena + s.io.off : s.R->text;
            if (s.R->text_bytes) memcpy(nt + o, src, s.R->text_bytes);
            const uint64_t from = s.R->text_borrowed ? s.io.off : 0;
            for (size_t i = 0; i < s.R->n_recs; i++) { s.R->recs[i].cigar_off += o - from; s.R->recs[i].md_off += o - from; }
            o += s.R->text_bytes;
        }
        g_result_pool_pinned.put(arena, arena_cap);
        arena = nt; arena_cap = cap; text_end = total;
    } else for (Sub &s : sub) if (s.R->text_bytes && s.io.off + s.R->text_bytes > text_end) text_end = s.io.off + s.R->text_bytes;
    {   /* every lane's records are ascending in chain index; a lane's chains are a subset of the call's: merge by global index */
        std::vector<uint32_t> first(n_chains + 1, 0);   /* records per chain -> where each chain's records start */
        for (Sub &s : sub) for (size_t i = 0; i < s.R->n_recs; i++) first[s.ids[s.R->recs[i].chain_id] + 1]++;
        for (size_t c = 0; c < n_chains; c++) first[c + 1] += first[c];
        for (Sub &s : sub) {
            for (size_t i = 0; i < s.R->n_recs; i++) { lf_sam_record r = s.R->recs[i]; r.chain_id = s.ids[r.chain_id]; R->recs[first[r.chain_id]++] = r; }
            const lf_chain_stats &a = s.R->stats;
            R->stats.round1_tasks += a.round1_tasks; R->stats.round2_extends += a.round2_extends; R->stats.round3_tasks += a.round3_tasks; R->stats.records += a.records;
            R->stats.ms_tasks = std::max(R->stats.ms_tasks, a.ms_tasks); R->stats.ms_round1 = std::max(R->stats.ms_round1, a.ms_round1);
            R->stats.ms_rounds23 = std::max(R->stats.ms_rounds23, a.ms_rounds23); R->stats.ms_emit = std::max(R->stats.ms_emit, a.ms_emit); R->stats.ms_merge = std::max(R->stats.ms_merge, a.ms_merge);
            delete s.R;
        }
    }
    R->n_recs = nrec;
    R->text = arena; R->text_cap = arena_cap; R->text_bytes = text_end;
    ctx->stats.kernel_launches = lfb_launches;
    *out = R;
    return LF_OK;
}

const lf_sam_record *lf_chain_results_records(const lf_chain_results *r, size_t *n) { if (n) *n = r ? r->n_recs : 0; return r ? r->recs : nullptr; }
const char *lf_chain_results_text(const lf_chain_results *r, size_t *bytes) { if (bytes) *bytes = r ? r->text_bytes : 0; return r ? r->text : nullptr; }
int lf_chain_results_stats(const lf_chain_results *r, lf_chain_stats *out) { if (!r || !out) return LF_ERR_BAD_ARG; *out = r->stats; return LF_OK; }
void lf_chain_results_free(lf_chain_results *r) { delete r; }

} /* extern "C" */
