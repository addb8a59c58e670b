/*
 * lordfast_gpu_glue.cpp -- the reference-side binding of liblfgpu.so: lordFAST with its alignment
 * stage on the GPU and everything else (CLI, index, seeding, windows, chaining, SAM writer) stock.
 *
 * This translation unit REPLACES the reference's LordFAST.o at link time (integration/Makefile).  It
 * pulls the reference's src/LordFAST.cpp in verbatim at build time (from /root/reference; nothing is
 * copied into this repository) with its chunk driver `mapSeqMT` (src/LordFAST.cpp:305-316) renamed,
 * and defines a batched `mapSeqMT` of its own, so the stock main() (src/baseFAST.cpp:32-84) calls it:
 *
 *   phase 0  ONE lf_gpu_seed_batch call for the chunk: getLocs_extend_whole_step (src/BWT.cpp:312-394) of every read,
 *            on bwa's index as the program loaded it (LF_GPU_SEED=0 keeps the stock per-read call in phase 1).
 *   phase 1  host threads, stock front-end per read: the seed lists of phase 0, findTopWins_coarse /
 *            findTopWins_fine, the coarse / fine decision of mapSeq (:535-568), and per window the
 *            seed selection + chain_seeds_n2 | chain_seeds_clasp of alignWin (:994-1059 / :1090-1144).
 *            Where alignWin would call the hook `alignChain` (:107, :1066 / :1151) the chain is
 *            copied out instead.
 *   phase 2  ONE lf_gpu_align_chains call for the chunk (include/lf_gpu.h): every alignChain_edlib
 *            (:1765-2258) of the chunk, batched on the GPU.
 *   phase 3  host threads, per read: the Sam_t lists are filled from the returned records, then
 *            alignWin's alnScore / totalScore arithmetic (:1067-1083 forward, :1152-1168 reverse --
 *            the reverse strand uses `gapPenalty`, the forward strand the literal 0.15), mapSeq's
 *            std::sort(compareSam) and the stock printSamEntry (:318-459).
 *
 * The output is the reference's SAM, record for record (tests/test_integration_sam.py compares the
 * sorted files; like the reference, the order of reads in the file depends on thread scheduling).
 */
#define mapSeqMT mapSeqMT_stock
#include "LordFAST.cpp" /* -I$(REF)/src : the reference's own file, read where it lies */
#undef mapSeqMT

#include <thread>
#include "bwa.h"
#include "lf_gpu.h"

extern bwaidx_t *_fmd_index; /* src/BWT.cpp:32 */
extern int32_t kCache;       /* src/BWT.cpp:34: k of the k-mer table the index was built with */

/* CUDA start-up (seconds on a cold box) overlaps the index load: this runs before main() of src/baseFAST.cpp */
__attribute__((constructor)) static void lfglue_prewarm(int argc, char **argv, char **)
{
    for (int i = 1; i < argc; i++) if (!strcmp(argv[i], "--search") || !strcmp(argv[i], "-S")) {
        /* the driver initialises every GPU it can see (seconds on an 8-GPU box): show it only the ones asked for */
        if (!getenv("CUDA_VISIBLE_DEVICES")) {
            const char *e = getenv("LF_GPU_DEVICES");
            setenv("CUDA_VISIBLE_DEVICES", e ? e : "0", 0);
            if (e) { /* the library then sees them renumbered 0..n-1 */
                std::string r; int n = 1; for (const char *p = e; *p; p++) n += *p == ',';
                for (int k = 0; k < n; k++) { if (k) r += ","; r += std::to_string(k); }
                setenv("LF_GPU_DEVICES", r.c_str(), 1);
            }
        }
        if (!getenv("LF_NO_PREWARM")) lf_gpu_prewarm();
        return;
    }
}

namespace lfglue {

enum ReadKind : uint8_t { RK_UNMAPPED = 0, RK_COARSE = 1, RK_FINE = 2 };

struct WinPlan {
    int64_t chain; /* index into the worker's chain list, -1 when chainLen <= 1 (:1086 / :1171) */
    uint8_t is_rev;
};
struct ReadPlan {
    uint32_t worker, first_win, n_win;
    uint8_t kind;
};
struct Worker { /* phase-1 output of one host thread */
    std::vector<lf_seed> seeds;
    std::vector<lf_chain> chains;
    std::vector<WinPlan> wins;
    uint64_t seed_base = 0, chain_base = 0; /* position of this worker's lists in the merged arrays */
};

static lf_gpu_ctx *g_ctx = nullptr;
static std::vector<Worker> g_workers;
static std::vector<ReadPlan> g_plan;
static std::vector<uint64_t> g_rec_first; /* per merged chain: first record, records of chain c are [first[c], first[c+1]) */
static const lf_sam_record *g_rec = nullptr;
static const char *g_text = nullptr;
static double g_ms[4];
static bool g_seed_ok = false;               /* lf_gpu_seed_init succeeded: chunks are seeded on the GPU */
static int g_ndev = 0;
static bool g_seed_gpu = false;              /* phase 0 ran: the seed lists of the chunk are in g_sd* */
static const lf_seed *g_sd[2] = {nullptr, nullptr};
static const uint64_t *g_sd_off[2] = {nullptr, nullptr};

static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void die(const char *what, int rc)
{
    fprintf(stderr, "[ERROR] (lordfast-gpu) %s failed: %d %s\n", what, rc, g_ctx ? lf_gpu_last_error(g_ctx) : "");
    exit(EXIT_FAILURE);
}

/* alignWin up to the alignChain call, for one window of the read worker `id` has seeded */
static void plan_window(Win_t &win, uint32_t rLen, uint32_t read_id, int id, Worker &W)
{
    const uint32_t margin = rLen >> 1;
    uint32_t chrBeg, chrEnd;
    bwt_get_chr_boundaries(win.tStart, win.tEnd, &chrBeg, &chrEnd);
    const int64_t lo = std::max<int64_t>((int64_t)win.tStart - (int64_t)margin, (int64_t)chrBeg);
    const int64_t hi = std::min<int64_t>((int64_t)win.tEnd + (int64_t)margin, (int64_t)chrEnd);
    const SeedList &src = win.isReverse ? _pf_seedsReverse[id] : _pf_seedsForward[id];
    SeedList &sel = _pf_seedsSelected[id];
    sel.num = 0;
    for (uint32_t i = 0; i < src.num; i++) {
        const int64_t p = (int64_t)src.list[i].tPos;
        if (p >= lo && p <= hi) sel.list[sel.num++] = src.list[i];
    }
    Chain_t &chain = _pf_topChains[id].list[0];
    if (chainAlg == CHAIN_ALG_CLASP) {
        const bool shift = lo > 2000000000; /* clasp works on 31-bit coordinates (:1031-1033) */
        if (shift) for (uint32_t i = 0; i < sel.num; i++) sel.list[i].tPos -= 2000000000;
        chain_seeds_clasp(sel.list, sel.num, chain);
        if (shift) for (uint32_t i = 0; i < chain.chainLen; i++) chain.seeds[i].tPos += 2000000000;
    } else if (chainAlg == CHAIN_ALG_DPN2) {
        chain_seeds_n2(sel.list, sel.num, chain);
    }
    WinPlan wp;
    wp.is_rev = win.isReverse ? 1 : 0;
    wp.chain = -1;
    if (chain.chainLen > 1) {
        lf_chain c;
        c.seed_off = W.seeds.size();
        c.n_seeds = chain.chainLen;
        c.read_id = read_id;
        c.is_rev = wp.is_rev;
        c.reserved = 0;
        for (uint32_t i = 0; i < chain.chainLen; i++) {
            lf_seed s = {chain.seeds[i].tPos, chain.seeds[i].qPos, chain.seeds[i].len};
            W.seeds.push_back(s);
        }
        wp.chain = (int64_t)W.chains.size();
        W.chains.push_back(c);
    }
    W.wins.push_back(wp);
}

/* phase 1: mapSeq (:461-580) without the alignment and without output */
static void *collect(void *idp)
{
    const int id = *(int *)idp;
    Worker &W = g_workers[id];
    int t;
    while ((t = pf_getNextRead()) < _pf_seqListSize) {
        Read *read = _pf_seqList + t;
        const uint32_t readLen = *read->length;
        ReadPlan &P = g_plan[t];
        P.worker = (uint32_t)id;
        P.first_win = (uint32_t)W.wins.size();
        P.n_win = 0;
        P.kind = RK_UNMAPPED;
        if (readLen < (uint32_t)MIN_READ_LEN) continue;
        if (g_seed_gpu) {   /* the lists lf_gpu_seed_batch wrote for read t, into the worker's SeedLists (Seed_t packs qPos and len) */
            SeedList *dst[2] = {_pf_seedsForward + id, _pf_seedsReverse + id};
            for (int k = 0; k < 2; k++) {
                const uint64_t a = g_sd_off[k][t], b = g_sd_off[k][t + 1];
                for (uint64_t i = a; i < b; i++) { Seed_t &o = dst[k]->list[i - a]; o.tPos = g_sd[k][i].tPos; o.qPos = g_sd[k][i].qPos; o.len = g_sd[k][i].len; }
                dst[k]->num = (uint32_t)(b - a);
            }
        } else getLocs_extend_whole_step(read->seq, readLen, SAMPLING_COUNT, _pf_seedsForward + id, _pf_seedsReverse + id);
        WinList &top = _pf_topWins[id];
        top.num = 0;
        findTopWins_coarse(readLen, _pf_seedsForward + id, 0, t + 1, id);
        findTopWins_coarse(readLen, _pf_seedsReverse + id, 1, -(t + 1), id);
        if (top.num == 0) continue;
        std::sort_heap(top.list, top.list + top.num, compareWin);
        const float scoreRatio = 4;
        if (top.list[0].score >= scoreRatio * top.list[1].score) {
            P.kind = RK_COARSE;
            plan_window(top.list[0], readLen, (uint32_t)t, id, W);
        } else {
            P.kind = RK_FINE;
            top.num = 0;
            const float minScore = (float)top.list[0].score / scoreRatio;
            findTopWins_fine(readLen, _pf_seedsForward + id, 0, t + _pf_seqListSize + 1, minScore, id);
            findTopWins_fine(readLen, _pf_seedsReverse + id, 1, -(t + _pf_seqListSize + 1), minScore, id);
            for (uint32_t i = 0; i < top.num; i++) plan_window(top.list[i], readLen, (uint32_t)t, id, W);
        }
        P.n_win = (uint32_t)W.wins.size() - P.first_win;
    }
    return NULL;
}

/* alignWin after the alignChain call (:1067-1090 / :1152-1175) */
static void finish_window(const Worker &W, const WinPlan &wp, uint32_t rLen, SamList_t &map)
{
    map.samList.clear();
    if (wp.chain < 0) { map.totalScore = -2 * rLen; return; }
    const uint64_t c = W.chain_base + (uint64_t)wp.chain;
    for (uint64_t k = g_rec_first[c]; k < g_rec_first[c + 1]; k++) {
        const lf_sam_record &r = g_rec[k];
        map.samList.emplace_back();
        Sam_t &s = map.samList.back();
        s.flag = (uint16_t)r.flag; s.pos = r.pos; s.posEnd = r.posEnd;
        s.qStart = r.qStart; s.qEnd = r.qEnd; s.nmCount = r.nmCount;
        s.cigar.assign(g_text + r.cigar_off, r.cigar_len);
        s.md.assign(g_text + r.md_off, r.md_len);
    }
    map.totalScore = 0;
    if (map.samList.empty()) return; /* the reference indexes an empty list here (:1073); never seen */
    uint32_t i;
    for (i = 0; i < map.samList.size(); i++) {
        map.samList[i].alnScore = map.samList[i].nmCount + (map.samList[i].qEnd - map.samList[i].qStart);
        map.totalScore += map.samList[i].nmCount;
    }
    for (i = 0; i < map.samList.size() - 1; i++) {
        uint32_t diff = abs((int64_t)map.samList[i + 1].pos - (int64_t)map.samList[i].posEnd) +
                        abs((int64_t)map.samList[i + 1].qStart - (int64_t)map.samList[i].qEnd);
        if (wp.is_rev) map.totalScore -= gapPenalty * diff; /* :1077 */
        else map.totalScore -= 0.15 * diff;                /* :1162 */
    }
    map.totalScore -= map.samList.front().qStart;
    map.totalScore -= (rLen - map.samList.back().qEnd);
}

/* phase 3: the rest of mapSeq -- score, sort, print */
static void *emit(void *idp)
{
    const int id = *(int *)idp;
    std::vector<char> seq_rev(SEQ_MAX_LENGTH), qual_rev(SEQ_MAX_LENGTH);
    std::ostringstream out;
    MapInfo &M = _pf_topMappings[id];
    int t;
    while ((t = pf_getNextRead()) < _pf_seqListSize) {
        Read *read = _pf_seqList + t;
        const uint32_t readLen = *read->length;
        const uint32_t qualLen = (*read->isFq ? *read->length : 1);
        const ReadPlan &P = g_plan[t];
        M.qName = read->name; M.seq = read->seq; M.qual = read->qual;
        if (P.kind == RK_UNMAPPED) {
            M.mappings[0].samList.clear();
            printSamEntry(M, readLen, 1, out);
            continue;
        }
        reverseComplement(read->seq, seq_rev.data(), *read->length);
        reverse(read->qual, qual_rev.data(), qualLen);
        M.seq_rev = seq_rev.data(); M.qual_rev = qual_rev.data();
        const Worker &W = g_workers[P.worker];
        for (uint32_t i = 0; i < P.n_win; i++) finish_window(W, W.wins[P.first_win + i], readLen, M.mappings[i]);
        if (P.kind == RK_FINE) std::sort(M.mappings, M.mappings + P.n_win, compareSam);
        printSamEntry(M, readLen, (int)P.n_win, out);
    }
    if (out.tellp() > 0) {
        pthread_mutex_lock(&_pf_outputLock);
        fprintf(_pf_outFile, "%s", out.str().c_str());
        pthread_mutex_unlock(&_pf_outputLock);
    }
    return NULL;
}

static void run_threads(void *(*fn)(void *))
{
    _pf_seqPos = 0;
    for (int i = 0; i < THREAD_COUNT; i++) pthread_create(_pf_threads + i, NULL, fn, THREAD_ID + i);
    for (int i = 0; i < THREAD_COUNT; i++) pthread_join(_pf_threads[i], NULL);
}

} // namespace lfglue

void mapSeqMT()
{
    using namespace lfglue;
    if (!g_ctx) {
        int ndev = 0, devs[16];
        if (const char *e = getenv("LF_GPU_DEVICES")) /* e.g. "0,1,2,3"; default: the current device */
            for (char *p = (char *)e; *p && ndev < 16;) { devs[ndev++] = (int)strtol(p, &p, 10); if (*p == ',') p++; }
        const double ti = now_ms();
        int rc = lf_gpu_init(&g_ctx, _fmd_index->pac, _fmd_index->bns->l_pac, ndev ? devs : nullptr, ndev);
        if (rc != LF_OK) die("lf_gpu_init", rc);
        fprintf(stderr, "[lordfast-gpu: lf_gpu_init %.1f ms] ", now_ms() - ti);
        g_ndev = ndev;
        const char *e = getenv("LF_GPU_SEED");
        if (!e || atoi(e) != 0) {   /* the FM index as bwa_idx_load and bwt_cache_load left it (src/BWT.cpp:190-224); the k-mer table is derived on the device */
            const bwt_t *b = _fmd_index->bwt;
            lf_fm_index fm;
            memset(&fm, 0, sizeof fm);
            fm.bwt = b->bwt; fm.bwt_size = b->bwt_size; fm.primary = b->primary;
            for (int i = 0; i < 5; i++) fm.L2[i] = b->L2[i];
            fm.seq_len = b->seq_len; fm.sa = b->sa; fm.n_sa = b->n_sa; fm.sa_intv = b->sa_intv;
            fm.k_cache = kCache; fm.cache = nullptr; fm.l_pac = _fmd_index->bns->l_pac;
            const double ts = now_ms();
            rc = lf_gpu_seed_init(g_ctx, &fm);
            if (rc != LF_OK) die("lf_gpu_seed_init", rc);
            g_seed_ok = true;
            fprintf(stderr, "[lordfast-gpu: lf_gpu_seed_init %.1f ms] ", now_ms() - ts);
        }
    }
    const double t0 = now_ms();
    /* gather the chunk's reads (each its own malloc block, src/Reads.cpp:84-90) into pinned memory kept between chunks */
    std::vector<uint64_t> off(_pf_seqListSize + 1, 0);
    for (int r = 0; r < _pf_seqListSize; r++) off[r + 1] = off[r] + *_pf_seqList[r].length;
    static uint8_t *bases = nullptr;   /* pinned, kept between chunks (pinning 100 MB costs tens of ms) */
    static size_t bases_cap = 0;
    if (off[_pf_seqListSize] + 1 > bases_cap) {
        if (bases) lf_gpu_host_free(bases);
        bases_cap = (off[_pf_seqListSize] + 1) * 5 / 4;
        bases = (uint8_t *)lf_gpu_host_alloc(bases_cap);
        if (!bases) die("lf_gpu_host_alloc", LF_ERR_NOMEM);
    }
    {   /* gather with all host threads */
        std::vector<std::thread> th;
        const int nt = THREAD_COUNT > 1 ? THREAD_COUNT : 1;
        for (int k = 0; k < nt; k++) th.emplace_back([&, k] {
            for (int r = (int)((int64_t)_pf_seqListSize * k / nt); r < (int)((int64_t)_pf_seqListSize * (k + 1) / nt); r++)
                memcpy(bases + off[r], _pf_seqList[r].seq, *_pf_seqList[r].length);
        });
        for (auto &t : th) t.join();
    }
    lf_reads rd = {bases, off.data(), (uint32_t)_pf_seqListSize};
    /* phase 0: seeding of the whole chunk on the GPU */
    lf_seed_results *sres = nullptr;
    g_seed_gpu = false;
    if (g_seed_ok && _pf_seqListSize > 0) {
        const lf_seed_params sp = {MIN_ANCHOR_LEN, SAMPLING_COUNT, MAX_REF_HITS};
        int rc = lf_gpu_seed_batch(g_ctx, &rd, &sp, &sres);
        if (rc != LF_OK) die("lf_gpu_seed_batch", rc);
        for (int k = 0; k < 2; k++) g_sd[k] = lf_seed_results_list(sres, k, &g_sd_off[k], nullptr);
        g_seed_gpu = true;
    }
    const double t0s = now_ms();
    g_workers.assign(THREAD_COUNT, Worker());
    g_plan.assign(_pf_seqListSize, ReadPlan());
    run_threads(collect);
    if (sres) lf_seed_results_free(sres);
    const double t1 = now_ms();

    /* merge the workers' lists */
    uint64_t ns = 0, nc = 0;
    for (Worker &W : g_workers) { W.seed_base = ns; W.chain_base = nc; ns += W.seeds.size(); nc += W.chains.size(); }
    /* seeds and chains of the chunk in pinned memory kept between chunks: the library then reads them over PCIe as they are
     * (from pageable memory it would first stage them, one more pass over 12 B per seed) */
    static lf_seed *seeds = nullptr; static lf_chain *chains = nullptr;
    static size_t seeds_cap = 0, chains_cap = 0;
    if (ns + 2 > seeds_cap) { if (seeds) lf_gpu_host_free(seeds); seeds_cap = (ns + 2) * 5 / 4; seeds = (lf_seed *)lf_gpu_host_alloc(seeds_cap * sizeof(lf_seed)); if (!seeds) die("lf_gpu_host_alloc", LF_ERR_NOMEM); }
    if (nc + 2 > chains_cap) { if (chains) lf_gpu_host_free(chains); chains_cap = (nc + 2) * 5 / 4; chains = (lf_chain *)lf_gpu_host_alloc(chains_cap * sizeof(lf_chain)); if (!chains) die("lf_gpu_host_alloc", LF_ERR_NOMEM); }
    for (Worker &W : g_workers) {
        if (!W.seeds.empty()) memcpy(&seeds[W.seed_base], W.seeds.data(), W.seeds.size() * sizeof(lf_seed));
        for (size_t k = 0; k < W.chains.size(); k++) { lf_chain c = W.chains[k]; c.seed_off += W.seed_base; chains[W.chain_base + k] = c; }
    }
    if (g_seed_gpu && g_ndev <= 1 && !getenv("LF_CHAIN_LANES")) rd.bases = nullptr;   /* the reads are still on the device from phase 0 */
    const bntseq_t *bns = _fmd_index->bns;
    std::vector<int64_t> coff(bns->n_seqs);
    std::vector<int32_t> clen(bns->n_seqs);
    for (int i = 0; i < bns->n_seqs; i++) { coff[i] = bns->anns[i].offset; clen[i] = bns->anns[i].len; }
    lf_contigs cg = {coff.data(), clen.data(), bns->n_seqs};

    lf_chain_results *res = nullptr;
    size_t nrec = 0;
    g_rec = nullptr; g_text = nullptr;
    if (nc) {
        int rc = lf_gpu_align_chains(g_ctx, &rd, &cg, seeds, chains, (size_t)nc, _fmd_index->pac, &res);
        if (rc != LF_OK) die("lf_gpu_align_chains", rc);
        g_rec = lf_chain_results_records(res, &nrec);
        g_text = lf_chain_results_text(res, nullptr);
    }
    g_rec_first.assign(nc + 1, 0);
    for (size_t k = 0; k < nrec; k++) g_rec_first[g_rec[k].chain_id + 1]++;
    for (uint64_t c = 0; c < nc; c++) g_rec_first[c + 1] += g_rec_first[c];
    const double t2 = now_ms();

    run_threads(emit);
    if (res) lf_chain_results_free(res);
    const double t3 = now_ms();
    g_ms[0] += t1 - t0; g_ms[1] += t2 - t1; g_ms[2] += t3 - t2;
    fprintf(stderr, "[lordfast-gpu: %llu chains, gather + GPU seeding %.1f ms, front-end %.1f ms, GPU alignment stage %.1f ms, scoring+SAM %.1f ms] ",
            (unsigned long long)nc, t0s - t0, t1 - t0s, t2 - t1, t3 - t2);
}
