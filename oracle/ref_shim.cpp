/*
 * ref_shim.cpp -- thin C-ABI access to the REFERENCE's own code, for tests and the CPU baseline.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled (oracle/Makefile, target `ref`) together with objects built
 * straight from /root/reference into oracle/_ref/libref_shim.so and oracle/_ref/lordfast_chaindump;
 * no reference source is copied into this repository.  The shim only *calls* the reference:
 *   edlibAlign / edlibFreeAlignResult   lib/edlib/edlib.h:172-192
 *   ksw_extend2                         lib/bwa/ksw.h:107-108
 *   alignChain_edlib                    src/LordFAST.cpp:1765 (through the reference's own hook,
 *                                       the global function pointer `alignChain`, :107)
 *   bwt_load/initRead/initializeFAST/readChunk/initFASTChunk/mapSeqMT/finalizeFAST
 *                                       src/baseFAST.cpp:32-84 (the stock search loop)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <string>
#include <vector>

#include "Common.h"
#include "CommandLineParser.h"
#include "BWT.h"
#include "LordFAST.h"
#include "bwa.h"
#include "edlib.h"
extern "C" {
#include "ksw.h"
}
#include "lf_oracle.h" /* lfo_seed / lfo_sam record layouts shared with the restatement */

/* globals and functions the reference defines without declaring them in a header */
extern bwaidx_t *_fmd_index;                                   /* src/BWT.cpp:32 */
extern uint8_t _pf_char2int[128];                              /* src/LordFAST.cpp:74 */
extern int8_t _pf_kswMatrix[25], _pf_kswMatrix_clip[25];       /* :75-76 */
extern MapInfo *_pf_topMappings;                               /* :62 */
extern void (*alignChain)(Chain_t &, char *, int32_t, int, SamList_t &); /* :107 */
void alignChain_edlib(Chain_t &chain, char *query, int32_t readLen, int isRev, SamList_t &map);

extern "C" {

int ref_align(const char *q, int ql, const char *t, int tl, int mode, int want_path, int *ed,
              int *endloc, unsigned char *ops, int *nops)
{
    EdlibAlignResult r = edlibAlign(q, ql, t, tl, edlibNewAlignConfig(-1, mode == 1 ? EDLIB_MODE_SHW : EDLIB_MODE_NW,
                                                                       want_path ? EDLIB_TASK_PATH : EDLIB_TASK_DISTANCE));
    *ed = r.editDistance;
    *endloc = r.endLocations ? r.endLocations[0] : -2;
    *nops = 0;
    if (want_path && r.alignment) {
        memcpy(ops, r.alignment, (size_t)r.alignmentLength);
        *nops = r.alignmentLength;
    }
    edlibFreeAlignResult(r);
    return 0;
}

int ref_extend(int qlen, const uint8_t *q, int tlen, const uint8_t *t, int m, const int8_t *mat, int o_del,
               int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, int *qle, int *tle)
{
    return ksw_extend2(qlen, q, tlen, t, m, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, qle, tle, 0, 0, 0);
}

/* ---- a fabricated in-memory index so that bwt_str_pac2char / bwt_get_chr_boundaries work ---- */
static bwaidx_t g_idx;
static bntseq_t g_bns;
static std::vector<bntann1_t> g_anns;
static std::vector<std::string> g_names;

static void init_tables()
{ /* what initializeFAST sets up for the alignment stage (src/LordFAST.cpp:158-187) */
    memset(_pf_char2int, 4, 128);
    _pf_char2int['A'] = _pf_char2int['a'] = 0; _pf_char2int['C'] = _pf_char2int['c'] = 1;
    _pf_char2int['G'] = _pf_char2int['g'] = 2; _pf_char2int['T'] = _pf_char2int['t'] = 3;
    int k = 0;
    for (int i = 0; i < 4; i++) { for (int j = 0; j < 4; j++) _pf_kswMatrix_clip[k++] = (i == j ? 2 : -16); _pf_kswMatrix_clip[k++] = 0; }
    for (int j = 0; j < 5; j++) _pf_kswMatrix_clip[k++] = 0;
    k = 0;
    for (int i = 0; i < 4; i++) { for (int j = 0; j < 4; j++) _pf_kswMatrix[k++] = (i == j ? 2 : -5); _pf_kswMatrix[k++] = 0; }
    for (int j = 0; j < 5; j++) _pf_kswMatrix[k++] = 0;
}

void ref_set_index(const uint8_t *pac, int64_t l_pac, int n_contigs, const int64_t *off, const int32_t *len)
{
    g_anns.assign((size_t)n_contigs, bntann1_t());
    g_names.resize((size_t)n_contigs);
    for (int i = 0; i < n_contigs; i++) {
        g_names[i] = "chr" + std::to_string(i + 1);
        g_anns[i].offset = off[i]; g_anns[i].len = len[i];
        g_anns[i].name = (char *)g_names[i].c_str(); g_anns[i].anno = (char *)"";
    }
    memset(&g_bns, 0, sizeof g_bns);
    g_bns.l_pac = l_pac; g_bns.n_seqs = n_contigs; g_bns.anns = g_anns.data();
    memset(&g_idx, 0, sizeof g_idx);
    g_idx.bns = &g_bns; g_idx.pac = (uint8_t *)pac;
    _fmd_index = &g_idx;
    init_tables();
}

/* ---- FM-index seeding: the reference's own index build / load and getLocs_extend_whole_step ---- */
extern bwtCache_t *_fmd_cacheTable;                              /* src/BWT.cpp:33 */
extern int32_t kCache;                                           /* src/BWT.cpp:34 */
static bwaidx_t *g_fm = NULL;

/* fasta: a FASTA file in a scratch directory; k_cache: the k of the k-mer table the index is built with (12 in the program) */
int ref_fm_load(const char *fasta, int k_cache)
{
    kCache = k_cache;
    if (bwt_load((char *)fasta)) return 1;   /* src/BWT.cpp:190: builds the index files next to the FASTA if they are missing */
    g_fm = _fmd_index;
    return 0;
}
int ref_fm_info(uint64_t *scalars /* primary, L2[5], seq_len, bwt_size, n_sa, sa_intv, l_pac, k_cache: 12 values */,
                const uint32_t **bwt, const uint64_t **sa, const void **cache)
{
    if (!g_fm) return 1;
    const bwt_t *b = g_fm->bwt;
    scalars[0] = b->primary; for (int i = 0; i < 5; i++) scalars[1 + i] = b->L2[i];
    scalars[6] = b->seq_len; scalars[7] = b->bwt_size; scalars[8] = b->n_sa; scalars[9] = (uint64_t)b->sa_intv;
    scalars[10] = (uint64_t)g_fm->bns->l_pac; scalars[11] = (uint64_t)kCache;
    *bwt = b->bwt; *sa = b->sa; *cache = _fmd_cacheTable;
    return 0;
}
/* one read through getLocs_extend_whole_step (src/BWT.cpp:312-394); q is NUL-terminated as Reads.cpp leaves it */
int ref_fm_seed(const char *q, uint32_t qlen, uint32_t hash_count, int min_anchor_len, int max_ref_hits,
                lfo_seed *fwd, int *n_fwd, lfo_seed *rev, int *n_rev)
{
    if (!g_fm) return 1;
    _fmd_index = g_fm;
    MIN_ANCHOR_LEN = min_anchor_len; MAX_REF_HITS = max_ref_hits;
    SeedList f, r;
    std::vector<Seed_t> fl((size_t)hash_count * max_ref_hits + 1), rl((size_t)hash_count * max_ref_hits + 1);   /* src/LordFAST.cpp:136-137 */
    f.list = fl.data(); r.list = rl.data(); f.num = r.num = 0;
    getLocs_extend_whole_step((char *)q, qlen, hash_count, &f, &r);
    for (uint32_t i = 0; i < f.num; i++) { fwd[i].tPos = fl[i].tPos; fwd[i].qPos = fl[i].qPos; fwd[i].len = fl[i].len; }
    for (uint32_t i = 0; i < r.num; i++) { rev[i].tPos = rl[i].tPos; rev[i].qPos = rl[i].qPos; rev[i].len = rl[i].len; }
    *n_fwd = (int)f.num; *n_rev = (int)r.num;
    return 0;
}
/* a batch on `threads` host threads (the program's own parallelism is one read per thread, src/LordFAST.cpp:295-316);
 * returns seconds; counts only (the lists are dropped) */
/* order-sensitive 64-bit digest of a seed list: sum over i of (i + 1) * mix(seed i); tests compute the same from the GPU's lists */
static inline uint64_t fm_seed_mix(uint32_t tPos, uint32_t qPos, uint32_t len)
{
    return (uint64_t)tPos * 0x9E3779B97F4A7C15ull + (uint64_t)qPos * 0xC2B2AE3D27D4EB4Full + (uint64_t)len * 0x165667B19E3779F9ull + 0x27D4EB2F165667C5ull;
}
static uint64_t fm_list_digest(const Seed_t *l, uint32_t n)
{
    uint64_t h = 0;
    for (uint32_t i = 0; i < n; i++) h += (uint64_t)(i + 1) * fm_seed_mix(l[i].tPos, l[i].qPos, l[i].len);
    return h;
}
struct FmBatchArg { const char *const *q; const uint32_t *qlen; int lo, hi; uint32_t hash_count, cap; uint64_t hits; uint64_t *dig; uint32_t *cnt; };
static void *fm_batch_worker(void *p)
{
    FmBatchArg *a = (FmBatchArg *)p;
    SeedList f, r;
    std::vector<Seed_t> fl(a->cap), rl(a->cap);
    f.list = fl.data(); r.list = rl.data();
    for (int i = a->lo; i < a->hi; i++) {
        getLocs_extend_whole_step((char *)a->q[i], a->qlen[i], a->hash_count, &f, &r);
        a->hits += f.num + r.num;
        if (a->dig) { a->dig[2 * i] = fm_list_digest(f.list, f.num); a->dig[2 * i + 1] = fm_list_digest(r.list, r.num); a->cnt[2 * i] = f.num; a->cnt[2 * i + 1] = r.num; }
    }
    return NULL;
}
/* a batch on `threads` host threads (the program's own parallelism is one read per thread, src/LordFAST.cpp:295-316);
 * returns seconds.  digests / counts (2 per read: forward, reverse) may be NULL: then only the total is kept. */
double ref_fm_seed_batch_digest(int n, const char *const *q, const uint32_t *qlen, uint32_t hash_count, int min_anchor_len, int max_ref_hits, int threads, uint64_t *hits,
                                uint64_t *digests, uint32_t *counts)
{
    if (!g_fm) return -1;
    _fmd_index = g_fm;
    MIN_ANCHOR_LEN = min_anchor_len; MAX_REF_HITS = max_ref_hits;
    std::vector<FmBatchArg> args((size_t)threads);
    std::vector<pthread_t> th((size_t)threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        args[t].q = q; args[t].qlen = qlen; args[t].lo = (int)((long long)n * t / threads); args[t].hi = (int)((long long)n * (t + 1) / threads);
        args[t].hash_count = hash_count; args[t].cap = hash_count * (uint32_t)max_ref_hits + 1; args[t].hits = 0; args[t].dig = digests; args[t].cnt = counts;
        pthread_create(&th[t], NULL, fm_batch_worker, &args[t]);
    }
    uint64_t tot = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); tot += args[t].hits; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (hits) *hits = tot;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
double ref_fm_seed_batch(int n, const char *const *q, const uint32_t *qlen, uint32_t hash_count, int min_anchor_len, int max_ref_hits, int threads, uint64_t *hits)
{
    return ref_fm_seed_batch_digest(n, q, qlen, hash_count, min_anchor_len, max_ref_hits, threads, hits, NULL, NULL);
}

static void fill_chain(Chain_t &c, std::vector<Seed_t> &store, const lfo_seed *seeds, int n)
{
    store.resize((size_t)n);
    for (int i = 0; i < n; i++) { store[i].tPos = seeds[i].tPos; store[i].qPos = seeds[i].qPos; store[i].len = seeds[i].len; }
    c.seeds = store.data(); c.chainLen = (uint32_t)n; c.score = 0;
}

static void *run_on_big_stack(void *(*fn)(void *), void *arg)
{ /* alignChain_edlib keeps ~2 MB of arrays on its stack (src/LordFAST.cpp:1770-1781) */
    pthread_attr_t a; pthread_t th; void *ret = NULL;
    pthread_attr_init(&a); pthread_attr_setstacksize(&a, 32u << 20);
    int rc = pthread_create(&th, &a, fn, arg);
    if (rc != 0) { fprintf(stderr, "[ref_shim] pthread_create failed (%d); running inline\n", rc); ret = fn(arg); }
    else pthread_join(th, &ret);
    pthread_attr_destroy(&a);
    return ret;
}

struct chain_call { const lfo_seed *seeds; int n; const char *query; int readLen, isRev; lfo_sam *out; int cap; int *n_out; };

static void *chain_call_fn(void *p)
{
    chain_call *a = (chain_call *)p;
    Chain_t c; std::vector<Seed_t> store; SamList_t map;
    fill_chain(c, store, a->seeds, a->n);
    alignChain_edlib(c, (char *)a->query, a->readLen, a->isRev, map);
    int k = 0;
    for (size_t i = 0; i < map.samList.size() && k < a->cap; i++, k++) {
        const Sam_t &s = map.samList[i];
        a->out[k].flag = s.flag; a->out[k].pos = s.pos; a->out[k].posEnd = s.posEnd;
        a->out[k].qStart = s.qStart; a->out[k].qEnd = s.qEnd; a->out[k].nmCount = s.nmCount;
        a->out[k].cigar = strdup(s.cigar.c_str()); a->out[k].md = strdup(s.md.c_str());
    }
    *a->n_out = k;
    return NULL;
}

int ref_align_chain(const lfo_seed *seeds, int n, const char *query, int readLen, int isRev, lfo_sam *out, int cap, int *n_out)
{
    chain_call a = { seeds, n, query, readLen, isRev, out, cap, n_out };
    run_on_big_stack(chain_call_fn, &a);
    return 0;
}

/* ---- CPU baseline: the reference's alignChain_edlib replayed over pre-dumped chains ---- */
struct replay_shared {
    int n_chains; const int64_t *seed_off; const lfo_seed *seeds; const char *const *query; const int32_t *readLen;
    const uint8_t *isRev; int next; pthread_mutex_t lock; int64_t n_sam;
};

static void *replay_worker(void *p)
{
    replay_shared *s = (replay_shared *)p;
    std::vector<Seed_t> store; int64_t ns = 0;
    for (;;) {
        pthread_mutex_lock(&s->lock); int i = s->next++; pthread_mutex_unlock(&s->lock);
        if (i >= s->n_chains) break;
        Chain_t c; SamList_t map;
        fill_chain(c, store, s->seeds + s->seed_off[i], (int)(s->seed_off[i + 1] - s->seed_off[i]));
        alignChain_edlib(c, (char *)s->query[i], s->readLen[i], s->isRev[i], map);
        ns += (int64_t)map.samList.size();
    }
    pthread_mutex_lock(&s->lock); s->n_sam += ns; pthread_mutex_unlock(&s->lock);
    return NULL;
}

/* returns wall seconds for one pass over all chains on `nthreads` threads */
double ref_replay_chains(int n_chains, const int64_t *seed_off, const lfo_seed *seeds, const char *const *query,
                         const int32_t *readLen, const uint8_t *isRev, int nthreads, int64_t *n_sam)
{
    replay_shared s = { n_chains, seed_off, seeds, query, readLen, isRev, 0, PTHREAD_MUTEX_INITIALIZER, 0 };
    std::vector<pthread_t> th((size_t)nthreads);
    pthread_attr_t a; pthread_attr_init(&a); pthread_attr_setstacksize(&a, 32u << 20);
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], &a, replay_worker, &s);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_attr_destroy(&a);
    if (n_sam) *n_sam = s.n_sam;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

} /* extern "C" */

/* ------------------------------------------------------------------------------------------ */
/* lordfast_chaindump: the stock search loop with the reference's `alignChain` hook pointed at  */
/* a recorder that forwards to alignChain_edlib and logs each chain and the records it yields.  */
/* ------------------------------------------------------------------------------------------ */
#ifdef LF_CHAINDUMP_MAIN
#include <zlib.h>
static FILE *g_dump = NULL;
static pthread_mutex_t g_dump_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_dump_hash = 0;      /* LF_CHAIN_DUMP_HASH=1: per record the lengths and zlib CRC-32 of CIGAR and MD instead of the strings */
static long g_dump_full = 0;     /* ... but the strings too for reads named r<k> with k < LF_CHAIN_DUMP_FULL */

static void dump_hook(Chain_t &chain, char *query, int32_t readLen, int isRev, SamList_t &map)
{
    size_t before = map.samList.size();
    alignChain_edlib(chain, query, readLen, isRev, map);
    const char *name = "?";
    for (int i = 0; i < THREAD_COUNT; i++)
        if (_pf_topMappings[i].seq == query || _pf_topMappings[i].seq_rev == query) name = _pf_topMappings[i].qName;
    pthread_mutex_lock(&g_dump_lock);
    fprintf(g_dump, "C\t%s\t%d\t%d\t%u\t", name, readLen, isRev, chain.chainLen);
    for (uint32_t i = 0; i < chain.chainLen; i++)
        fprintf(g_dump, "%u,%u,%u;", chain.seeds[i].tPos, (unsigned)chain.seeds[i].qPos, (unsigned)chain.seeds[i].len);
    fprintf(g_dump, "\n");
    const bool full = !g_dump_hash || (name[0] == 'r' && atol(name + 1) < g_dump_full);
    for (size_t i = before; i < map.samList.size(); i++) {
        const Sam_t &s = map.samList[i];
        if (g_dump_hash)
            fprintf(g_dump, "H\t%u\t%u\t%u\t%u\t%u\t%d\t%zu\t%lu\t%zu\t%lu\n", (unsigned)s.flag, s.pos, s.posEnd, s.qStart, s.qEnd, s.nmCount,
                    s.cigar.size(), (unsigned long)crc32(0L, (const Bytef *)s.cigar.data(), (uInt)s.cigar.size()),
                    s.md.size(), (unsigned long)crc32(0L, (const Bytef *)s.md.data(), (uInt)s.md.size()));
        if (full)
            fprintf(g_dump, "S\t%u\t%u\t%u\t%u\t%u\t%d\t%s\t%s\n", (unsigned)s.flag, s.pos, s.posEnd, s.qStart, s.qEnd, s.nmCount,
                    s.cigar.c_str(), s.md.c_str());
    }
    pthread_mutex_unlock(&g_dump_lock);
}

int main(int argc, char *argv[])
{
    const char *path = getenv("LF_CHAIN_DUMP");
    g_dump_hash = getenv("LF_CHAIN_DUMP_HASH") && atoi(getenv("LF_CHAIN_DUMP_HASH"));
    g_dump_full = getenv("LF_CHAIN_DUMP_FULL") ? atol(getenv("LF_CHAIN_DUMP_FULL")) : 0;
    if (parseCommandLine(argc, argv)) return EXIT_FAILURE;
    if (indexingMode) return bwt_index(refFile) ? EXIT_FAILURE : 0;
    g_dump = fopen(path ? path : "chains.txt", "w");
    if (!g_dump) { perror("LF_CHAIN_DUMP"); return EXIT_FAILURE; }
    Read *seqList; unsigned int seqListSize;
    if (bwt_load(refFile)) return EXIT_FAILURE;
    if (!initRead(seqFile, 100000000)) return EXIT_FAILURE;
    initializeFAST();
    alignChain = &dump_hook;
    while (readChunk(&seqList, &seqListSize) > 0) {
        initFASTChunk(seqList, seqListSize);
        mapSeqMT();
        releaseChunk();
    }
    finalizeFAST();
    fclose(g_dump);
    return 0;
}
#endif
