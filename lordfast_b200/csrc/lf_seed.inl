/*
 * lf_seed.inl -- FM-index seeding on the GPU: lf_gpu_seed_init / lf_gpu_seed_batch of
 * include/lf_gpu.h.  Included at the end of lf_pipeline.inl (product build: nvcc; test-only build: the fiber emulator).
 *
 * What it computes is getLocs_extend_whole_step (src/BWT.cpp:312-394) for every read of a batch, over bwa's FM index of
 * reference + reverse complement as the reference program loads it (bwt_t, lib/bwa/bwt.h:44-57; the k-mer table of
 * src/BWT.cpp:60-138).  The reference's loop is sequential in three places, and each is restated so that the result is
 * the same list in the same order:
 *   sample positions   seed_pos += step in double precision, truncated: one thread per read repeats exactly those adds
 *                      (k_seed_positions); everything after that is parallel over (read, sample).
 *   longest match      the reference tries lengths MIN_ANCHOR_LEN, +1, +2 ... with a fresh backward search each, until
 *                      one fails.  k_seed_probe runs the first search for every sample, k_seed_extend lengthens the
 *                      matches that exist, one thread each, picking the lengths by doubling and bisection ("length L
 *                      occurs" is monotone): the interval of a length is a property of the index, not of the order of
 *                      the probes, and a search is a handful of dependent 64-byte reads, which is what a GPU hides
 *                      best with many threads in flight.
 *   containment filter "kept if pos + m > last_pos" where last_pos is the end of the last kept sample: kept samples have
 *                      strictly increasing ends and a dropped one ends at or before last_pos, so last_pos is the running
 *                      maximum of the ends of all earlier samples with an acceptable hit count -- a prefix maximum, one
 *                      warp per read (k_seed_filter).
 * The hits of the kept samples are then located one thread per hit (k_seed_locate: bwt_sa, lib/bwa/bwt.c:86-98) and
 * split stably into the forward and reverse lists (exclusive scan of the strand flags + k_seed_scatter).
 *
 * HBM layout: bwt (64-byte blocks: four 64-bit counts + 128 bases), sa (64-bit, one per sa_intv suffixes), the k-mer
 * table (16 bytes per k-mer), per sample {pos, m, sp, hit count}, per hit a 12-byte seed + strand flag.  Everything is
 * random 32/64-byte access: the bound is HBM/L2 latency x threads in flight, not bandwidth.
 */
#pragma once

struct LfFmDev {
    const uint32_t *bwt; const unsigned long long *sa; const lf_fm_cache_entry *cache;
    unsigned long long primary, L2[5], seq_len, n_sa, sa_mask;
    int sa_shift, k_cache;
    long long l_pac;
};

/* nst_nt4_table (lib/bwa/bntseq.c:47-64) as far as seeding can tell: 0..3 for ACGTacgt, 4 for everything else */
__device__ __forceinline__ uint32_t lf_fm_nt4(uint32_t ch)
{
    const uint32_t uc = ch & 0xdfu, code = ((uc >> 1) ^ (uc >> 2)) & 3u;
    return ((0x54474341u >> (8u * code)) & 0xffu) == uc ? code : 4u;
}

__device__ __forceinline__ unsigned long long lf_fm_L2(const LfFmDev &fm, uint32_t c)
{ /* L2[c], c = 0..4, without indexing the kernel parameter dynamically (that would copy it to local memory) */
    return c == 0 ? fm.L2[0] : c == 1 ? fm.L2[1] : c == 2 ? fm.L2[2] : c == 3 ? fm.L2[3] : fm.L2[4];
}

/* One 64-byte block of bwt_t::bwt: four 64-bit running counts, then 128 bases in eight words, first base on top. */
struct LfFmBlock { uint4 c01, c23, b0, b1; };
__device__ __forceinline__ LfFmBlock lf_fm_block(const LfFmDev &fm, unsigned long long k)
{
    const uint4 *p = (const uint4 *)(fm.bwt + ((k >> 7) << 4));
    LfFmBlock b;
    b.c01 = __ldg(p); b.c23 = __ldg(p + 1); b.b0 = __ldg(p + 2); b.b1 = __ldg(p + 3);
    return b;
}
/* bases equal to c among the first nb (0..16) of a word: what __occ_aux (lib/bwa/bwt.c:100-108) counts, with the mask
 * applied after the comparison (so the correction for masked-out bases that look like A, :129, is not needed) */
__device__ __forceinline__ uint32_t lf_fm_word_count(uint32_t x, uint32_t m1, uint32_t m2, int nb)
{
    const uint32_t t = ((x ^ m2) >> 1) & (x ^ m1) & 0x55555555u;
    const int n = nb < 0 ? 0 : nb > 16 ? 16 : nb;
    return (uint32_t)__popc(t & (uint32_t)(0xffffffff00000000ull >> (2 * n)));
}
/* occurrences of c in rows 0..k of the block that holds k (k already without the $ row) */
__device__ __forceinline__ unsigned long long lf_fm_block_occ(const LfFmBlock &b, unsigned long long k, uint32_t c)
{
    const uint32_t lo = c == 0 ? b.c01.x : c == 1 ? b.c01.z : c == 2 ? b.c23.x : b.c23.z;
    const uint32_t hi = c == 0 ? b.c01.y : c == 1 ? b.c01.w : c == 2 ? b.c23.y : b.c23.w;
    const uint32_t m1 = (c & 1u) ? 0u : 0xffffffffu, m2 = (c & 2u) ? 0u : 0xffffffffu;
    const int kk = (int)(k & 127ull) + 1;      /* bases of the block to count */
    uint32_t n = lf_fm_word_count(b.b0.x, m1, m2, kk) + lf_fm_word_count(b.b0.y, m1, m2, kk - 16) + lf_fm_word_count(b.b0.z, m1, m2, kk - 32) + lf_fm_word_count(b.b0.w, m1, m2, kk - 48)
               + lf_fm_word_count(b.b1.x, m1, m2, kk - 64) + lf_fm_word_count(b.b1.y, m1, m2, kk - 80) + lf_fm_word_count(b.b1.z, m1, m2, kk - 96) + lf_fm_word_count(b.b1.w, m1, m2, kk - 112);
    return ((unsigned long long)hi << 32 | lo) + n;
}

__device__ __forceinline__ unsigned long long lf_fm_occ(const LfFmDev &fm, unsigned long long k, uint32_t c)
{ /* bwt_occ, lib/bwa/bwt.c:110-132 */
    if (k == fm.seq_len) return lf_fm_L2(fm, c + 1u) - lf_fm_L2(fm, c);
    if (k == ~0ull) return 0;
    k -= (k >= fm.primary) ? 1ull : 0ull;
    return lf_fm_block_occ(lf_fm_block(fm, k), k, c);
}

/* one backward-search step: the interval [k, l] of a pattern becomes that of c + pattern (src/BWT.cpp:287-291).
 * As in bwt_2occ (lib/bwa/bwt.c:135-166) the two lookups share the block when both rows lie in it -- the usual case once
 * the interval is small. */
__device__ __forceinline__ void lf_fm_step(const LfFmDev &fm, unsigned long long &k, unsigned long long &l, uint32_t c)
{
    unsigned long long ok, ol;
    const unsigned long long k1 = k - 1ull;
    const unsigned long long ka = k1 - ((k1 >= fm.primary) ? 1ull : 0ull), la = l - ((l >= fm.primary) ? 1ull : 0ull);
    if (k1 != ~0ull && l != fm.seq_len && k1 != fm.seq_len && (ka >> 7) == (la >> 7)) {
        const LfFmBlock b = lf_fm_block(fm, ka);
        ok = lf_fm_block_occ(b, ka, c); ol = lf_fm_block_occ(b, la, c);
    } else { ok = lf_fm_occ(fm, k1, c); ol = lf_fm_occ(fm, l, c); }
    const unsigned long long base = lf_fm_L2(fm, c);
    k = base + ok + 1ull;
    l = base + ol;
}

__device__ __forceinline__ unsigned long long lf_fm_inv_psi(const LfFmDev &fm, unsigned long long k)
{ /* bwt_invPsi, lib/bwa/bwt.c:52-58 */
    const unsigned long long x = k - ((k > fm.primary) ? 1ull : 0ull);
    const uint32_t c = (__ldg(fm.bwt + ((x >> 7) << 4) + 8 + ((x & 127ull) >> 4)) >> ((~x & 15ull) << 1)) & 3u;
    const unsigned long long r = lf_fm_L2(fm, c) + lf_fm_occ(fm, k, c);
    return k == fm.primary ? 0ull : r;
}

/* sum over the warp, added once to a global counter (work statistics of a batch) */
__device__ __forceinline__ void lf_seed_count(unsigned long long *ctr, uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(LF_FULL, v, o);
    if ((threadIdx.x & 31u) == 0 && v) atomicAdd(ctr, (unsigned long long)v);
}

/* bwt_count_exact_cached (src/BWT.cpp:265-298) on the bases s[0 .. len): the last k_cache of them index the k-mer table,
 * the others extend to the left one by one.  Positions at or past `avail` read as the NUL the reference finds there.
 * Returns the number of occurrences (0: none, k and l untouched as in the reference). */
__device__ __forceinline__ long long lf_fm_count(const LfFmDev &fm, const uint8_t *__restrict__ s, long long avail, int len, unsigned long long &k_out, unsigned long long &l_out, uint32_t &steps)
{
    if ((long long)len > avail) return 0;          /* a base past the end of the read is not ACGT */
    uint32_t idx = 0;
    for (int i = len - 1; i >= len - fm.k_cache; --i) {
        const uint32_t c = lf_fm_nt4(__ldg(s + i));
        if (c > 3u) return 0;
        idx = idx * 4u + c;
    }
    const lf_fm_cache_entry e = fm.cache[idx];
    if (e.beg > e.end) return 0;
    unsigned long long k = e.beg, l = e.end;
    for (int i = len - fm.k_cache - 1; i >= 0; --i) {
        const uint32_t c = lf_fm_nt4(__ldg(s + i));
        if (c > 3u) return 0;
        lf_fm_step(fm, k, l, c);
        steps++;
        if (k > l) return 0;
    }
    k_out = k; l_out = l;
    return (long long)(l - k + 1ull);
}

/* The k-mer table bwt_cache_gen writes (src/BWT.cpp:60-138), one thread per k-mer: the digits of the index are the
 * characters of the backward search, first character = most significant digit; an interval that has become empty is
 * handed down unchanged (:89-99). */
__global__ void __launch_bounds__(256) k_seed_cache_build(LfFmDev fm, lf_fm_cache_entry *out, uint32_t n)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    unsigned long long k = 0, l = fm.seq_len;
    for (int d = fm.k_cache - 1; d >= 0; --d) {
        if (k > l) break;
        lf_fm_step(fm, k, l, (idx >> (2 * d)) & 3u);
    }
    lf_fm_cache_entry e; e.beg = k; e.end = l;
    out[idx] = e;
}

/* seed_pos += step in double, truncated to uint32 (src/BWT.cpp:320-322, :389-390): one thread per read */
__global__ void __launch_bounds__(128) k_seed_positions(const uint64_t *__restrict__ read_off, uint32_t n_reads, uint32_t S, uint32_t *__restrict__ pos)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t qlen = (uint32_t)(read_off[r + 1] - read_off[r]);
    const double step = (double)qlen / (double)S;
    double sp = 0;
    uint32_t spi = 0;
    uint32_t *o = pos + (size_t)r * S;
    for (uint32_t i = 0; i < S; i++) {
        o[i] = spi;
#if defined(__CUDA_ARCH__)
        sp = __dadd_rn(sp, step);   /* a plain add, never contracted */
#else
        sp += step;
#endif
        spi = (uint32_t)sp;
    }
}

/* Longest match at one sample (src/BWT.cpp:328-342), in two kernels so that the lanes of a warp have the same kind of work:
 * k_seed_probe   one thread per (read, sample): the first search (MIN_ANCHOR_LEN bases).  Five samples in six of a noisy
 *                read end here; the others are appended to a dense list.
 * k_seed_extend  one thread per listed sample: lengthen the match.  The reference adds one base at a time until a search
 *                fails.  "Length L occurs" is monotone in L (a longer pattern contains the shorter one, a bad base or the
 *                end of the read stays inside it), so the same last success is found by doubling the increment until a
 *                search fails and bisecting the gap: about 2 (m - 12) search steps instead of (m - 12)^2 / 2, and the
 *                interval kept is the one of the last successful search. */
__global__ void __launch_bounds__(128) k_seed_probe(LfFmDev fm, const uint8_t *__restrict__ bases, const uint64_t *__restrict__ read_off, uint32_t n_reads, uint32_t S,
                                                    int min_len, long long max_hits, const uint32_t *__restrict__ pos, uint32_t *__restrict__ mlen,
                                                    unsigned long long *__restrict__ sp_out, uint32_t *__restrict__ cnt, uint32_t *__restrict__ list, uint32_t *__restrict__ n_list,
                                                    unsigned long long *__restrict__ ctr)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t steps = 0;
    bool more = false;
    if (g < (size_t)n_reads * S) {
        const uint32_t r = (uint32_t)(g / S);
        const uint64_t ro = read_off[r];
        const uint32_t p = pos[g];
        unsigned long long sp = 0, ep = 0;
        const long long occ = lf_fm_count(fm, bases + ro + p, (long long)(read_off[r + 1] - ro) - (long long)p, min_len, sp, ep, steps);
        mlen[g] = (uint32_t)min_len;
        sp_out[g] = sp;
        cnt[g] = (occ > 0 && occ < max_hits) ? (uint32_t)occ : 0u;   /* final if the match cannot be lengthened */
        more = occ > 0;
    }
    /* dense list of the samples that matched: one atomic per warp */
    const uint32_t bal = __ballot_sync(LF_FULL, more), lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == 0 && bal) base = atomicAdd(n_list, (uint32_t)__popc(bal));
    base = __shfl_sync(LF_FULL, base, 0);
    if (more) list[base + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = (uint32_t)g;
    lf_seed_count(ctr, steps);   /* backward-search steps: two occurrence lookups each */
}

/* Lanes of a warp would otherwise wait for the one whose match is longest (5 of 32 lanes busy on a noisy-read chunk).
 * Every lane instead runs a small state machine -- take a sample from the list, start a search (k-mer table), take ONE
 * backward-search step, decide the next length -- so that each pass of the loop is the same step for all lanes that are in
 * the middle of a search, and a lane that has finished its sample takes the next one at once. */
#define LF_SEED_NONE 0xffffffffu
__device__ __forceinline__ uint32_t lf_seed_take(uint32_t *next, bool need)
{ /* index of the next work item for the lanes that need one (one atomic per warp) */
    const uint32_t bal = __ballot_sync(LF_FULL, need), lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (bal == 0u) return LF_SEED_NONE;
    if (lane == (uint32_t)(__ffs((int)bal) - 1)) base = atomicAdd(next, (uint32_t)__popc(bal));
    base = __shfl_sync(LF_FULL, base, __ffs((int)bal) - 1);
    return need ? base + (uint32_t)__popc(bal & ((1u << lane) - 1u)) : LF_SEED_NONE;
}

__global__ void __launch_bounds__(128) k_seed_extend(LfFmDev fm, const uint8_t *__restrict__ bases, const uint64_t *__restrict__ read_off, uint32_t S,
                                                     long long max_hits, const uint32_t *__restrict__ pos, uint32_t *__restrict__ mlen,
                                                     unsigned long long *__restrict__ sp_out, uint32_t *__restrict__ cnt, const uint32_t *__restrict__ list, const uint32_t *__restrict__ n_list,
                                                     uint32_t *__restrict__ next, unsigned long long *__restrict__ ctr)
{
    const uint32_t n = *n_list;
    uint32_t g = LF_SEED_NONE, steps = 0;
    const uint8_t *s = bases;
    long long avail = 0, occ = -1;
    int m = 0, inc = 1, hi = 0, L = 0, i = 0;
    unsigned long long k = 0, l = 0, sp = 0;
    bool searching = false;
    for (;;) {
        const uint32_t item = lf_seed_take(next, g == LF_SEED_NONE);
        if (item != LF_SEED_NONE && item < n) {
            g = list[item];
            const uint32_t r = g / S;
            const uint64_t ro = read_off[r];
            const uint32_t p = pos[g];
            s = bases + ro + p;
            avail = (long long)(read_off[r + 1] - ro) - (long long)p;
            m = (int)mlen[g]; sp = sp_out[g]; occ = -1; inc = 1; hi = 0; searching = false;   /* occ -1: the first search's count is already in cnt[g] */
        }
        if (__all_sync(LF_FULL, g == LF_SEED_NONE)) break;
        if (g == LF_SEED_NONE) continue;
        if (searching) {   /* one step of the current search */
            const uint32_t c = lf_fm_nt4(__ldg(s + i));
            bool fail = c > 3u;
            if (!fail) { lf_fm_step(fm, k, l, c); steps++; fail = k > l; }
            if (fail) { hi = L; searching = false; }
            else if (--i < 0) { occ = (long long)(l - k + 1ull); sp = k; m = L; if (!hi) inc <<= 1; searching = false; }
        } else if (hi && hi - m <= 1) {   /* the last success is final */
            if (occ >= 0) { mlen[g] = (uint32_t)m; sp_out[g] = sp; cnt[g] = occ < max_hits ? (uint32_t)occ : 0u; }
            g = LF_SEED_NONE;
        } else {   /* next length: doubling until a search fails, then bisection; start the search at the k-mer table */
            L = hi ? (m + hi) >> 1 : m + inc;
            bool ok = (long long)L <= avail;
            uint32_t idx = 0;
            for (int j = L - 1; ok && j >= L - fm.k_cache; --j) {
                const uint32_t c = lf_fm_nt4(__ldg(s + j));
                if (c > 3u) ok = false;
                idx = idx * 4u + c;
            }
            if (ok) { const lf_fm_cache_entry e = fm.cache[idx]; k = e.beg; l = e.end; ok = k <= l; }
            i = L - fm.k_cache - 1;
            if (!ok) hi = L;
            else if (i < 0) { occ = (long long)(l - k + 1ull); sp = k; m = L; if (!hi) inc <<= 1; }
            else searching = true;
        }
    }
    lf_seed_count(ctr, steps);
}

/* containment filter (src/BWT.cpp:345, :387): cnt[i] stays only where pos + m exceeds the ends of all earlier
 * acceptable samples of the read.  One warp per read. */
__global__ void __launch_bounds__(128) k_seed_filter(uint32_t n_reads, uint32_t S, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ mlen, uint32_t *__restrict__ cnt)
{
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (r >= n_reads) return;
    uint32_t last = 0;
    for (uint32_t i0 = 0; i0 < S; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const size_t g = (size_t)r * S + i;
        const uint32_t c = i < S ? cnt[g] : 0u;
        const uint32_t e = c ? pos[g] + mlen[g] : 0u;
        uint32_t inc = e;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(LF_FULL, inc, o); if ((int)lane >= o) inc = inc > v ? inc : v; }
        uint32_t exc = __shfl_up_sync(LF_FULL, inc, 1);
        if (lane == 0) exc = 0;
        exc = exc > last ? exc : last;
        if (i < S && c && !(e > exc)) cnt[g] = 0u;
        const uint32_t tot = __shfl_sync(LF_FULL, inc, 31);
        last = last > tot ? last : tot;
    }
}

/* locate (src/BWT.cpp:348-384; bwt_sa, lib/bwa/bwt.c:86-98).  Hit h of the batch belongs to the sample whose exclusive hit
 * offset is the last one <= h (found read first, then sample: both searches stay in cache).  The walk to the next sampled
 * row takes a geometrically distributed number of inverse-psi steps (mean sa_intv, maximum of a warp ~4x that), so the
 * lanes take hits from a queue and every pass of the loop is one step for all of them. */
__global__ void __launch_bounds__(128) k_seed_locate(LfFmDev fm, const uint64_t *__restrict__ read_off, uint32_t n_reads, uint32_t S, const unsigned long long *__restrict__ hoff /* n_reads*S + 1 */,
                                                     const uint32_t *__restrict__ pos, const uint32_t *__restrict__ mlen, const unsigned long long *__restrict__ sp,
                                                     uint32_t n_hits, lf_seed *__restrict__ hits, uint32_t *__restrict__ is_rev, uint32_t *__restrict__ next, unsigned long long *__restrict__ ctr)
{
    uint32_t h = LF_SEED_NONE, g = 0, steps = 0;
    unsigned long long k = 0, sa = 0;
    for (;;) {
        const uint32_t item = lf_seed_take(next, h == LF_SEED_NONE);
        if (item != LF_SEED_NONE && item < n_hits) {
            h = item;
            uint32_t lo = 0, hi = n_reads;                 /* last read r with hoff[r*S] <= h */
            while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (hoff[(size_t)mid * S] <= h) lo = mid; else hi = mid; }
            const unsigned long long *ho = hoff + (size_t)lo * S;
            const uint32_t r = lo;
            lo = 0; hi = S;                                /* last sample i with ho[i] <= h (empty samples share an offset with their successor: take the last) */
            while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (ho[mid] <= h) lo = mid; else hi = mid; }
            g = r * S + lo;
            k = sp[g] + (h - ho[lo]);
            sa = 0;
        }
        if (__all_sync(LF_FULL, h == LF_SEED_NONE)) break;
        if (h == LF_SEED_NONE) continue;
        if (k & fm.sa_mask) { ++sa; k = lf_fm_inv_psi(fm, k); steps++; continue; }
        unsigned long long sapos = sa + __ldg(fm.sa + (k >> fm.sa_shift));
        const uint32_t r = g / S, m = mlen[g], p = pos[g];
        const uint32_t qlen = (uint32_t)(read_off[r + 1] - read_off[r]);
        lf_seed sd;
        uint32_t rev = 0;
        if (sapos >= (unsigned long long)fm.l_pac) { /* reverse strand */
            sapos = ((unsigned long long)fm.l_pac << 1) - sapos - m;
            sd.qPos = (qlen - p - m) & 0xfffffu;
            rev = 1;
        } else sd.qPos = p & 0xfffffu;
        sd.tPos = (uint32_t)sapos;
        sd.len = m & 0xfffu;                          /* Seed_t: qPos is a 20-bit and len a 12-bit field (src/LordFAST.h:30-35) */
        hits[h] = sd;
        is_rev[h] = rev;
        h = LF_SEED_NONE;
    }
    lf_seed_count(ctr, steps);   /* inverse-psi steps: one base + one occurrence lookup each */
}

/* stable split into the two lists: rbefore[h] = reverse hits before h */
__global__ void __launch_bounds__(256) k_seed_scatter(const lf_seed *__restrict__ hits, const uint32_t *__restrict__ is_rev, const unsigned long long *__restrict__ rbefore,
                                                      unsigned long long n_hits, lf_seed *__restrict__ fwd, lf_seed *__restrict__ rev)
{
    const unsigned long long h = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hits) return;
    const unsigned long long rb = rbefore[h];
    if (is_rev[h]) rev[rb] = hits[h]; else fwd[h - rb] = hits[h];
}
__global__ void __launch_bounds__(256) k_seed_read_offsets(uint32_t n_reads, uint32_t S, const unsigned long long *__restrict__ hoff, const unsigned long long *__restrict__ rbefore /* n_hits + 1 */,
                                                           unsigned long long *__restrict__ fwd_off, unsigned long long *__restrict__ rev_off)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    const unsigned long long base = hoff[(size_t)r * S];
    const unsigned long long rb = rbefore[base];
    rev_off[r] = rb; fwd_off[r] = base - rb;
}

/* ------------------------------------------------------------------------------------------ */
/* host side                                                                                   */
/* ------------------------------------------------------------------------------------------ */
struct lf_seed_results {
    uint32_t n_reads = 0;
    const lf_seed *fwd = nullptr, *rev = nullptr;
    const uint64_t *fwd_off = nullptr, *rev_off = nullptr;
    size_t n_fwd = 0, n_rev = 0;
};

namespace {

struct PinnedBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t need) { if (need <= cap) return 0; lfb_host_free(p); p = nullptr; cap = 0; const size_t want = need + need / 4 + 4096; p = lfb_host_alloc(want); if (!p) return -5; cap = want; return 0; }
    void release() { lfb_host_free(p); p = nullptr; cap = 0; }
};

struct SeedState {
    LfFmDev fm = {};
    bool ready = false;
    size_t n_cache = 0;
    LfbBuf bwt, sa, cache, pos, mlen, sp, cnt, hoff, hits, is_rev, rbefore, fwd, rev, fwd_off, rev_off, ctr, list;
    PinnedBuf h_fwd, h_rev, h_off, h_tot;
    lf_seed_stats st = {};
#ifndef LF_EMU
    cudaEvent_t ev[4] = {};
#endif
};

void seed_state_free_fn(void *p)
{
    SeedState *s = (SeedState *)p;
    LfbBuf *bufs[] = { &s->bwt, &s->sa, &s->cache, &s->pos, &s->mlen, &s->sp, &s->cnt, &s->hoff, &s->hits, &s->is_rev, &s->rbefore, &s->fwd, &s->rev, &s->fwd_off, &s->rev_off, &s->ctr, &s->list };
    for (LfbBuf *b : bufs) b->release();
    s->h_fwd.release(); s->h_rev.release(); s->h_off.release(); s->h_tot.release();
#ifndef LF_EMU
    for (int k = 0; k < 4; k++) if (s->ev[k]) cudaEventDestroy(s->ev[k]);
#endif
    delete s;
}

SeedState &seed_state(lf_gpu_ctx *ctx)
{
    if (!ctx->seed_state) { ctx->seed_state = new SeedState(); ctx->seed_state_free = seed_state_free_fn; }
    return *(SeedState *)ctx->seed_state;
}

} // namespace

extern "C" {

int lf_gpu_seed_init(lf_gpu_ctx *ctx, const lf_fm_index *fm)
{
    if (!ctx || !fm || !fm->bwt || !fm->sa || fm->bwt_size == 0 || fm->n_sa == 0) return ctx ? fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_seed_init: null index") : LF_ERR_BAD_ARG;
    if (fm->sa_intv <= 0 || (fm->sa_intv & (fm->sa_intv - 1)) || fm->k_cache < 1 || fm->k_cache > 14 || fm->l_pac <= 0 || fm->seq_len == 0)
        return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_seed_init: sa_intv must be a power of two, 1 <= k_cache <= 14");
    if (fm->bwt_size < (((fm->seq_len - 1) >> 7) + 1) * 16 - 8 || fm->n_sa < (fm->seq_len >> 0) / (uint64_t)fm->sa_intv)
        return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_seed_init: bwt / sa shorter than seq_len asks for");
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    SeedState &S = seed_state(ctx);
    S.ready = false;
    const size_t bwt_bytes = (size_t)fm->bwt_size * 4, sa_bytes = (size_t)fm->n_sa * 8;
    S.n_cache = (size_t)1 << (2 * fm->k_cache);
    if (S.bwt.reserve(bwt_bytes + 64) || S.sa.reserve(sa_bytes) || S.cache.reserve(S.n_cache * sizeof(lf_fm_cache_entry))) return fail(ctx, LF_ERR_NOMEM, "seed index");
    LF_TRY(lfb_memset((uint8_t *)S.bwt.p + bwt_bytes, 0, 64, d.stream));   /* the last block may be read to its end */
    LF_TRY(lfb_h2d(S.bwt.p, fm->bwt, bwt_bytes, d.stream));
    LF_TRY(lfb_h2d(S.sa.p, fm->sa, sa_bytes, d.stream));
    LfFmDev v;
    v.bwt = S.bwt.as<uint32_t>(); v.sa = S.sa.as<unsigned long long>(); v.cache = S.cache.as<lf_fm_cache_entry>();
    v.primary = fm->primary; for (int i = 0; i < 5; i++) v.L2[i] = fm->L2[i];
    v.seq_len = fm->seq_len; v.n_sa = fm->n_sa; v.sa_mask = (unsigned long long)fm->sa_intv - 1ull;
    v.sa_shift = 0; while ((1 << v.sa_shift) < fm->sa_intv) v.sa_shift++;
    v.k_cache = fm->k_cache; v.l_pac = fm->l_pac;
    S.fm = v;
    if (fm->cache) LF_TRY(lfb_h2d(S.cache.p, fm->cache, S.n_cache * sizeof(lf_fm_cache_entry), d.stream));
    else LFB_LAUNCH(k_seed_cache_build, (unsigned)((S.n_cache + 255) / 256), 256, 0, d.stream, v, S.cache.as<lf_fm_cache_entry>(), (uint32_t)S.n_cache);
    LF_TRY(lfb_last_error());
    LF_TRY(lfb_sync(d.stream));
#ifndef LF_EMU
    for (int k = 0; k < 4; k++) if (!S.ev[k]) cudaEventCreate(&S.ev[k]);
#endif
    S.ready = true;
    return LF_OK;
}

int lf_gpu_seed_cache_download(lf_gpu_ctx *ctx, lf_fm_cache_entry *out, size_t n)
{
    if (!ctx || !out) return LF_ERR_BAD_ARG;
    SeedState &S = seed_state(ctx);
    if (!S.ready || n != S.n_cache) return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_seed_cache_download: no index, or n != 4^k_cache");
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    LF_TRY(lfb_d2h(out, S.cache.p, n * sizeof(lf_fm_cache_entry), d.stream));
    LF_TRY(lfb_sync(d.stream));
    return LF_OK;
}

int lf_gpu_seed_batch(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_seed_params *prm, lf_seed_results **out)
{
    if (!ctx || !prm || !out) return LF_ERR_BAD_ARG;
    *out = nullptr;
    SeedState &S = seed_state(ctx);
    if (!S.ready) return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_seed_batch: lf_gpu_seed_init first");
    if (prm->sampling_count <= 0 || prm->max_ref_hits <= 0 || prm->min_anchor_len < S.fm.k_cache || prm->min_anchor_len > 4095)
        return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_seed_batch: sampling_count, max_ref_hits > 0 and k_cache <= min_anchor_len <= 4095");
    if (reads) { const int rc = lf_gpu_upload_reads(ctx, reads); if (rc) return rc; }
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    const uint32_t n_reads = d.n_reads, SC = (uint32_t)prm->sampling_count;
    lf_seed_results *res = new lf_seed_results();
    res->n_reads = n_reads;
    const size_t off_bytes = (size_t)(n_reads + 1) * 8;
    if (S.h_off.reserve(2 * off_bytes) || S.h_tot.reserve(64)) { delete res; return fail(ctx, LF_ERR_NOMEM, "seed offsets"); }
    uint64_t *h_fwd_off = (uint64_t *)S.h_off.p, *h_rev_off = h_fwd_off + (n_reads + 1);
    res->fwd_off = h_fwd_off; res->rev_off = h_rev_off;
    if (n_reads == 0) { h_fwd_off[0] = 0; h_rev_off[0] = 0; *out = res; return LF_OK; }
    const size_t np = (size_t)n_reads * SC;
    if (np >> 31) { delete res; return fail(ctx, LF_ERR_BAD_ARG, "lf_gpu_seed_batch: n_reads * sampling_count >= 2^31: split the batch"); }
    lfb_stream s = d.stream;
    auto bail = [&](int code, const char *msg) { delete res; return fail(ctx, code, msg); };
    if (S.pos.reserve(np * 4) || S.mlen.reserve(np * 4) || S.sp.reserve(np * 8) || S.cnt.reserve(np * 4) || S.hoff.reserve((np + 1) * 8) || S.ctr.reserve(32) || S.list.reserve(np * 4 + 16)) return bail(LF_ERR_NOMEM, "seed samples");
    if (lfb_memset(S.ctr.p, 0, 32, s)) return bail(LF_ERR_CUDA, "seed counters");
#ifndef LF_EMU
    cudaEventRecord(S.ev[0], s);
#endif
    LFB_LAUNCH(k_seed_positions, (n_reads + 127) / 128, 128, 0, s, d.read_off.as<uint64_t>(), n_reads, SC, S.pos.as<uint32_t>());
    LFB_LAUNCH(k_seed_probe, (unsigned)((np + 127) / 128), 128, 0, s, S.fm, d.bases.as<uint8_t>(), d.read_off.as<uint64_t>(), n_reads, SC, (int)prm->min_anchor_len, (long long)prm->max_ref_hits,
               S.pos.as<uint32_t>(), S.mlen.as<uint32_t>(), S.sp.as<unsigned long long>(), S.cnt.as<uint32_t>(), S.list.as<uint32_t>(), (uint32_t *)(S.ctr.as<unsigned long long>() + 2), S.ctr.as<unsigned long long>());
    {   /* persistent grid over the dense list (its length stays on the device) */
        const unsigned blocks = (unsigned)std::min<size_t>((np + 127) / 128, (size_t)148 * 16);
        LFB_LAUNCH(k_seed_extend, blocks, 128, 0, s, S.fm, d.bases.as<uint8_t>(), d.read_off.as<uint64_t>(), SC, (long long)prm->max_ref_hits,
                   S.pos.as<uint32_t>(), S.mlen.as<uint32_t>(), S.sp.as<unsigned long long>(), S.cnt.as<uint32_t>(), S.list.as<uint32_t>(), (const uint32_t *)(S.ctr.as<unsigned long long>() + 2), (uint32_t *)(S.ctr.as<unsigned long long>() + 3), S.ctr.as<unsigned long long>());
    }
    LFB_LAUNCH(k_seed_filter, (n_reads + 3) / 4, 128, 0, s, n_reads, SC, S.pos.as<uint32_t>(), S.mlen.as<uint32_t>(), S.cnt.as<uint32_t>());
    if (lfb_scan_excl_total(d.tmp, S.cnt.as<uint32_t>(), S.hoff.as<unsigned long long>(), np, s)) return bail(LF_ERR_CUDA, "seed scan");
#ifndef LF_EMU
    cudaEventRecord(S.ev[1], s);
#endif
    unsigned long long *h_tot = (unsigned long long *)S.h_tot.p;
    if (lfb_d2h(h_tot, S.hoff.as<unsigned long long>() + np, 8, s) || lfb_sync(s)) return bail(LF_ERR_CUDA, "seed totals");
    const unsigned long long H = h_tot[0];
    if (H >> 31) return bail(LF_ERR_BAD_ARG, "lf_gpu_seed_batch: 2^31 or more hits in one batch: split it");
    if (S.hits.reserve((size_t)H * sizeof(lf_seed) + 16) || S.is_rev.reserve((size_t)H * 4 + 16) || S.rbefore.reserve(((size_t)H + 1) * 8) ||
        S.fwd.reserve((size_t)H * sizeof(lf_seed) + 16) || S.rev.reserve((size_t)H * sizeof(lf_seed) + 16) || S.fwd_off.reserve(off_bytes) || S.rev_off.reserve(off_bytes))
        return bail(LF_ERR_NOMEM, "seed hits");
#ifndef LF_EMU
    cudaEventRecord(S.ev[2], s);
#endif
    if (H) {
        LFB_LAUNCH(k_seed_locate, (unsigned)std::min<size_t>((size_t)((H + 127) / 128), (size_t)148 * 16), 128, 0, s, S.fm, d.read_off.as<uint64_t>(), n_reads, SC, S.hoff.as<unsigned long long>(), S.pos.as<uint32_t>(), S.mlen.as<uint32_t>(),
                   S.sp.as<unsigned long long>(), (uint32_t)H, S.hits.as<lf_seed>(), S.is_rev.as<uint32_t>(), (uint32_t *)(S.ctr.as<unsigned long long>() + 3) + 1, S.ctr.as<unsigned long long>() + 1);
        if (lfb_scan_excl_total(d.tmp, S.is_rev.as<uint32_t>(), S.rbefore.as<unsigned long long>(), (size_t)H, s)) return bail(LF_ERR_CUDA, "seed strand scan");
        LFB_LAUNCH(k_seed_scatter, (unsigned)((H + 255) / 256), 256, 0, s, S.hits.as<lf_seed>(), S.is_rev.as<uint32_t>(), S.rbefore.as<unsigned long long>(), H, S.fwd.as<lf_seed>(), S.rev.as<lf_seed>());
    } else if (lfb_memset(S.rbefore.p, 0, 8, s)) return bail(LF_ERR_CUDA, "seed memset");
    LFB_LAUNCH(k_seed_read_offsets, (n_reads + 256) / 256, 256, 0, s, n_reads, SC, S.hoff.as<unsigned long long>(), S.rbefore.as<unsigned long long>(),
               S.fwd_off.as<unsigned long long>(), S.rev_off.as<unsigned long long>());
#ifndef LF_EMU
    cudaEventRecord(S.ev[3], s);
#endif
    if (lfb_last_error()) return bail(LF_ERR_CUDA, "seed kernels");
    if (lfb_d2h(h_fwd_off, S.fwd_off.p, off_bytes, s) || lfb_d2h(h_rev_off, S.rev_off.p, off_bytes, s) || lfb_d2h(h_tot + 1, S.rbefore.as<unsigned long long>() + H, 8, s) || lfb_d2h(h_tot + 2, S.ctr.p, 16, s) || lfb_sync(s))
        return bail(LF_ERR_CUDA, "seed offsets download");
    const unsigned long long n_rev = h_tot[1], n_fwd = H - n_rev;
    if (S.h_fwd.reserve((size_t)n_fwd * sizeof(lf_seed) + 16) || S.h_rev.reserve((size_t)n_rev * sizeof(lf_seed) + 16)) return bail(LF_ERR_NOMEM, "seed lists (pinned)");
    if ((n_fwd && lfb_d2h(S.h_fwd.p, S.fwd.p, (size_t)n_fwd * sizeof(lf_seed), s)) || (n_rev && lfb_d2h(S.h_rev.p, S.rev.p, (size_t)n_rev * sizeof(lf_seed), s)) || lfb_sync(s))
        return bail(LF_ERR_CUDA, "seed lists download");
    res->fwd = (const lf_seed *)S.h_fwd.p; res->rev = (const lf_seed *)S.h_rev.p;
    res->n_fwd = (size_t)n_fwd; res->n_rev = (size_t)n_rev;
    S.st.positions = np; S.st.hits = H; S.st.search_steps = h_tot[2]; S.st.locate_steps = h_tot[3];
#ifndef LF_EMU
    cudaEventElapsedTime(&S.st.search_ms, S.ev[0], S.ev[1]);
    cudaEventElapsedTime(&S.st.locate_ms, S.ev[2], S.ev[3]);
#endif
    *out = res;
    return LF_OK;
}

const lf_seed *lf_seed_results_list(const lf_seed_results *r, int reverse, const uint64_t **offsets, size_t *n)
{
    if (!r) return nullptr;
    if (offsets) *offsets = reverse ? r->rev_off : r->fwd_off;
    if (n) *n = reverse ? r->n_rev : r->n_fwd;
    return reverse ? r->rev : r->fwd;
}

void lf_seed_results_free(lf_seed_results *r) { delete r; }

int lf_gpu_seed_stats(lf_gpu_ctx *ctx, lf_seed_stats *out)
{
    if (!ctx || !out) return LF_ERR_BAD_ARG;
    *out = seed_state(ctx).st;
    return LF_OK;
}

} // extern "C"
