/*
 * lf_kernels.cuh -- device code of the batched alignment stage (sm_100a).
 *
 * Kernels (all integer / bit-vector work, no tensor cores -- nothing here is a contraction):
 *   k_pack_reads      reads (ASCII) -> three bit planes per read (lo, hi, not-ACGT), warp ballot.
 *                     Replaces edlib's transformSequences/buildPeq (lib/edlib/edlib.cpp:281, :1350):
 *                     Eq for a target symbol is two LOP3s on the planes.
 *   k_align_prep      per task: size class, sort key, op-slot and scratch sizes, validation.
 *   k_myers_small<NW> thread-per-task Myers/Hyyro bit-vector NW/SHW for q <= 32*NW rows held in
 *                     registers as one long word (carry chain instead of per-block hin/hout), with
 *                     checkpoints every 16 columns and a windowed recompute for the traceback.
 *                     Replaces edlibAlign for tasks below the 1 MiB rule (edlib.cpp:101-221,
 *                     :657-858 distance, :872-1071 traceback).
 *   k_myers_band<NB>  the same with the op planes of every column written to HBM and a walk over them instead
 *                     of the recompute (prefix-mode tasks of q <= 128; optionally a sliding band + retry).
 *   k_myers_bandreg<NB>  thread-per-task with a band of NB words in registers: the whole column for global-mode
 *                     tasks of q <= 128, a sliding band with a certificate (Ukkonen's argument, edlib.cpp:722-760)
 *                     for near-diagonal ones up to 512 rows; checkpoints + banded recompute, nothing per column in HBM.
 *   k_myers_large     warp-per-task wavefront (lane = 32-row word, lane-skewed columns, hin/hout
 *                     by __shfl_up) for everything else, including the Hirschberg recursion with
 *                     edlib's split rule (edlib.cpp:1090-1143, :1161-1330).
 *   k_ksw_extend      ksw_extend2 (lib/bwa/ksw.c:380-479) with its adaptive band, z-drop and
 *                     stale-cell behaviour kept exactly; int32 scores.
 *
 *   k_chain_tasks / k_chain_triggers / k_emit_slots   the chain operator's device side: round-1 task list from the
 *                     chains, clip / split trigger tests (src/LordFAST.cpp:1840, :1952, :2175), run-length CIGAR / MD /
 *                     record assembly (edlibCigar_*, edlibMD_*, :1570-1763).
 *
 * The results every kernel must reproduce are pure functions of the two strings (SURVEY.md
 * Appendix A/B): Levenshtein distance; for SHW the first target prefix (incl. the empty one, -1)
 * reaching the minimum; the path by canonical traceback (Up > Left > Diagonal from the bottom-right
 * corner) below the size rule 20*ceil(q/64)*t + 8*t < 2^20, else split at column t/2 on the
 * smallest row x in [1,q-1] with L[x]+R[x]==best (then x=0, then x=q).
 *
 * This header is compiled by nvcc into liblfgpu.so and, for debugging without a GPU, by g++ on top
 * of tests/emu/cuda_emu.h (test-only; never shipped).
 */
#pragma once
#include <stdint.h>
#include "lf_gpu.h"

#ifndef LF_EMU
#define LF_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char lf_dyn_smem_raw[]; type *name = (type *)lf_dyn_smem_raw
#else
#define LF_DYN_SMEM(type, name) EMU_DYN_SMEM(type, name)
#endif

#define LF_FULL 0xffffffffu
#define LF_K1_BLOCK 128   /* threads per block of k_myers_small */
#ifndef LF_K1_C
#define LF_K1_C 8         /* checkpoint interval in columns */
#endif
#define LF_NSMALL 8       /* register-resident size classes */
#define LF_CLS_LARGE 16   /* class id of k_myers_large tasks (small classes are 2*i + shw) */
#define LF_CLS_BAD 17
#define LF_CLS_BANDREG0 18 /* 18..21: global-mode tasks of size classes 4..7 that k_myers_bandreg runs in a sliding band */
/* 22..28: k_myers_group classes (LANES lanes per task).  GPnn: with path, leaves of up to nn words; GDnn: distance only */
#define LF_CLS_GP16 22    /*   9 ..  16 words: 4 lanes x 4 words (off-diagonal / prefix-mode tasks of 257 .. 512 rows) */
#define LF_CLS_GP32 23    /*  17 ..  32 words: 8 x 4 */
#define LF_CLS_GP64 24    /*  33 ..  64 words: 8 x 8 */
#define LF_CLS_GD32 25    /*  17 ..  32 words: 8 x 4 */
#define LF_CLS_GD64 26    /*  33 ..  64 words: 8 x 8 */
#define LF_CLS_GD128 27   /*  65 .. 128 words: 16 x 8 */
#define LF_CLS_GD256 28   /* 129 .. 256 words: 32 x 8 */
#define LF_NGROUPCLS 7
#define LF_NCLS 29
#ifndef LF_BUCKET_BITS
#define LF_BUCKET_BITS 5   /* target-length buckets of the task sort: 2^-LF_BUCKET_BITS octave wide (1/8 octave: 1.834 ms per config-2 step, 1/32: 1.808) */
#endif
#define LF_KEY_SHIFT (32 - 5 - 5 - LF_BUCKET_BITS)   /* sort keys use the top bits: 5 of class, 5 + LF_BUCKET_BITS of length bucket */
#define LF_LARGE_STACK 96 /* Hirschberg stack entries per warp (depth <= log2(t)+2) */
#define LF_CLIP_LEN 500   /* _pf_clipLen, src/LordFAST.cpp:88: heads / tails longer than this are first asked for their distance only */

__host__ __device__ __forceinline__ int lf_small_nw(int i)
{ /* words of 32 rows held in registers by size class i */
    return i == 0 ? 1 : i == 1 ? 2 : i == 2 ? 3 : i == 3 ? 4 : i == 4 ? 6 : i == 5 ? 8 : i == 6 ? 12 : 16;
}
__host__ __device__ __forceinline__ int lf_small_class(uint32_t nwords)
{
    return nwords <= 1 ? 0 : nwords <= 2 ? 1 : nwords <= 3 ? 2 : nwords <= 4 ? 3 : nwords <= 6 ? 4 : nwords <= 8 ? 5 : nwords <= 12 ? 6 : nwords <= 16 ? 7 : -1;
}
/* edlib's choice between full traceback and Hirschberg, keyed to 64-bit blocks (edlib.cpp:1117-1119) */
__host__ __device__ __forceinline__ bool lf_is_leaf(uint32_t q, uint32_t t)
{
    unsigned long long b64 = (q + 63u) / 64u;
    return 20ull * b64 * t + 8ull * t < (1ull << 20) || t < 2;
}
__host__ __device__ __forceinline__ uint64_t lf_plane_word_off(uint64_t read_byte_off, uint32_t r)
{ /* first plane word of read r: one zero pad word before and (at least) one after every read */
    return (read_byte_off >> 5) + 3ull * r + 1ull;
}
__host__ __device__ __forceinline__ uint32_t lf_k1_ckpt_bytes(uint32_t t, int nw)
{
    uint32_t nck = t ? (t - 1) / LF_K1_C : 0;
    return (nck * (uint32_t)nw * 8u + 15u) & ~15u;
}

__host__ __device__ __forceinline__ int lf_bandreg_nb(int sc) { return sc == 4 ? 3 : sc == 7 ? 5 : 4; } /* band words for q <= 192 | 256, 384 | 512 */
__host__ __device__ __forceinline__ bool lf_bandreg_eligible(uint32_t q, uint32_t t, int sc)
{ /* near-diagonal enough for the band certificate of k_myers_bandreg to have room for the usual 15 % distance */
    const int dlt = q > t ? (int)(q - t) : (int)(t - q);
    return sc >= 4 && 3 * dlt <= 32 * (lf_bandreg_nb(sc) - 1) - 7 && q / t < 32u;
}

/* Everything a kernel needs about one resident batch. */
struct LfDev {
    const uint8_t *pac; int64_t l_pac;
    const uint8_t *bases; const uint64_t *read_off; uint32_t n_reads;
    uint32_t *plo, *phi, *pnn;             /* read bit planes */
    const lf_align_task *tasks; uint32_t n_tasks;
    lf_align_result *res;
    uint32_t *ops;                         /* 2-bit op stream, 16 ops per word */
    const uint64_t *slot_end;              /* inclusive scan of per-task slot words */
    const uint64_t *scr_off;               /* exclusive scan of per-task scratch bytes (small classes) */
    uint8_t *scratch;
    uint8_t *planes;                       /* op planes of k_myers_band, one region per warp group */
    uint32_t bandreg;                      /* bit i: near-diagonal global tasks of size class 4+i (128 < q <= 512) go to k_myers_bandreg */
    uint32_t groupk;                       /* k_myers_group classes in use: bit 0 distance-only (GDnn), bit 1 GP32 / GP64, bit 2 GP16 */
};

struct LfCounters { /* written by k_align_prep, read back by the host (one small D2H per batch) */
    uint32_t hist[LF_NCLS];
    uint32_t max_q, max_t;
    uint32_t gmax_t[LF_NGROUPCLS + 1];     /* longest target per k_myers_group class (sizes the plane scratch) */
    unsigned long long max_planes, cells, word_columns, small_word_columns;
};

/* ------------------------------------------------------------------------------------------ */
/* sequence access                                                                            */
/* ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t lf_tsym(const uint8_t *__restrict__ pac, int64_t l)
{ /* _get_pac, src/BWT.cpp:310 */
    return ((uint32_t)__ldg(pac + (l >> 2)) >> ((~(uint32_t)l & 3u) << 1)) & 3u;
}

/* Sequential reader of the 2-bit reference: one aligned 32-bit load covers 16 columns and the next
 * word is requested one refill ahead, so the load latency (the per-column byte load was the top
 * long-scoreboard stall of the first version) is off the dependency chain. */
struct LfTCursor {
    const uint32_t *p32; int64_t widx; uint32_t buf, nxt; int left, dir;
    __device__ __forceinline__ static uint32_t be(uint32_t v)
    { /* bytes hold bases MSB-first: make base j of the word sit at bits 31-2j .. 30-2j */
        return __byte_perm(v, 0u, 0x0123u);
    }
    __device__ __forceinline__ void init(const uint8_t *pac, int64_t l, int d)
    {
        p32 = (const uint32_t *)pac; dir = d; widx = l >> 4;
        const int j = (int)(l & 15);
        buf = be(__ldg(p32 + widx));
        if (d > 0) { buf <<= 2 * j; left = 16 - j; widx++; }
        else { buf >>= 2 * (15 - j); left = j + 1; widx--; }
        nxt = be(__ldg(p32 + (widx < 0 ? 0 : widx)));
    }
    __device__ __forceinline__ void next_masks(uint32_t &slo, uint32_t &shi)
    { /* all-ones / all-zeros masks of the two bits of the next symbol */
        if (left == 0) { buf = nxt; left = 16; widx += dir; nxt = be(__ldg(p32 + (widx < 0 ? 0 : widx))); }
        if (dir > 0) { shi = (uint32_t)((int32_t)buf >> 31); slo = (uint32_t)((int32_t)(buf << 1) >> 31); buf <<= 2; }
        else { slo = 0u - (buf & 1u); shi = 0u - ((buf >> 1) & 1u); buf >>= 2; }
        left--;
    }
    __device__ __forceinline__ uint32_t next()
    {
        if (left == 0) { buf = nxt; left = 16; widx += dir; nxt = be(__ldg(p32 + (widx < 0 ? 0 : widx))); }
        uint32_t sym;
        if (dir > 0) { sym = buf >> 30; buf <<= 2; }
        else { sym = buf & 3u; buf >>= 2; }
        left--;
        return sym;
    }
};

/* The 2-bit reference as a stream in task order: the next 16 symbols are the top 32 bits of a 64-bit window
 * (first symbol at bits 31:30), whatever the strand; the word after next is requested one refill ahead. */
struct LfTStream {
    const uint32_t *p32; int64_t widx; uint32_t cur, nxt; int off, dir;
    __device__ __forceinline__ uint32_t norm(uint32_t v) const
    {
        v = __byte_perm(v, 0u, 0x0123u);   /* base j of the word at bits 31-2j .. 30-2j */
        if (dir < 0) { v = __brev(v); v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1); } /* base 15 first */
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t *pac, int64_t l, int d)
    {
        p32 = (const uint32_t *)pac; dir = d; widx = l >> 4;
        const int j = (int)(l & 15);
        cur = norm(__ldg(p32 + widx));
        off = d > 0 ? 2 * j : 2 * (15 - j);
        widx += d;
        nxt = norm(__ldg(p32 + (widx < 0 ? 0 : widx)));
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(nxt, cur, (uint32_t)off); }
    __device__ __forceinline__ void advance(int n)
    { /* n <= 16 symbols consumed */
        off += 2 * n;
        if (off >= 32) { off -= 32; cur = nxt; widx += dir; nxt = norm(__ldg(p32 + (widx < 0 ? 0 : widx))); }
    }
};

/* ... read one column at a time (k_myers_small): 16 symbols buffered, one test per column */
struct LfTStreamCol {
    LfTStream ts; uint32_t buf; int left;
    __device__ __forceinline__ void init(const uint8_t *pac, int64_t l, int d) { ts.init(pac, l, d); left = 0; buf = 0; }
    __device__ __forceinline__ void next_masks(uint32_t &slo, uint32_t &shi)
    { /* all-ones / all-zeros masks of the two bits of the next symbol */
        if (left == 0) { buf = ts.peek(); ts.advance(16); left = 16; }
        shi = (uint32_t)((int32_t)buf >> 31); slo = (uint32_t)((int32_t)(buf << 1) >> 31);
        buf <<= 2; left--;
    }
};

/* not-Eq of 32 query rows against one target symbol: two 3-input LOP3s (the compiler's own association of
 * the five inputs costs three) */
__device__ __forceinline__ uint32_t lf_neq(uint32_t qlo, uint32_t qhi, uint32_t qnn, uint32_t slo, uint32_t shi)
{
#if defined(__CUDA_ARCH__)
    uint32_t x, r;
    asm("lop3.b32 %0, %1, %2, %3, 0xbe;" : "=r"(x) : "r"(qlo), "r"(slo), "r"(qnn)); /* (a ^ b) | c */
    asm("lop3.b32 %0, %1, %2, %3, 0xf6;" : "=r"(r) : "r"(x), "r"(qhi), "r"(shi));   /* a | (b ^ c) */
    return r;
#else
    return (qlo ^ slo) | qnn | (qhi ^ shi);
#endif
}

struct LfQView { int64_t bit0; int dir; uint32_t comp; }; /* element k lives at plane bit bit0 + dir*k */
struct LfTView { int64_t t0; int dir; };                   /* element k is pac base t0 + dir*k */

__device__ __forceinline__ void lf_task_views(const LfDev &d, const lf_align_task &t, LfQView &qv, LfTView &tv)
{
    uint64_t ro = d.read_off[t.read_id];
    int64_t L = (int64_t)(d.read_off[t.read_id + 1] - ro);
    int rev1 = (t.flags & (LF_F_REVERSE_BOTH | LF_F_RC_QUERY)) != 0;
    int64_t oi0 = rev1 ? (int64_t)t.q_off + t.q_len - 1 : (int64_t)t.q_off;
    int odir = rev1 ? -1 : 1;
    uint32_t comp = (t.flags & LF_F_RC_QUERY) ? 1u : 0u;
    int64_t f0 = oi0;
    int dir = odir;
    if (t.flags & LF_F_READ_REV) { f0 = L - 1 - oi0; dir = -odir; comp ^= 1u; }
    qv.bit0 = (int64_t)lf_plane_word_off(ro, t.read_id) * 32 + f0;
    qv.dir = dir;
    qv.comp = comp ? 0xffffffffu : 0u;
    int revt = (t.flags & LF_F_REVERSE_BOTH) != 0;
    tv.t0 = revt ? (int64_t)t.t_off + t.t_len - 1 : (int64_t)t.t_off;
    tv.dir = revt ? -1 : 1;
}
__device__ __forceinline__ LfQView lf_qsub(const LfQView &v, int64_t off, int64_t len, bool reversed)
{
    LfQView r = v;
    if (!reversed) r.bit0 = v.bit0 + v.dir * off;
    else { r.bit0 = v.bit0 + v.dir * (off + len - 1); r.dir = -v.dir; }
    return r;
}
__device__ __forceinline__ LfTView lf_tsub(const LfTView &v, int64_t off, int64_t len, bool reversed)
{
    LfTView r = v;
    if (!reversed) r.t0 = v.t0 + v.dir * off;
    else { r.t0 = v.t0 + v.dir * (off + len - 1); r.dir = -v.dir; }
    return r;
}
__device__ __forceinline__ uint32_t lf_bits32(const uint32_t *__restrict__ p, int64_t B)
{ /* plane bits [B, B+32) */
    int64_t w = B >> 5;
    return __funnelshift_r(__ldg(p + w), __ldg(p + w + 1), (uint32_t)B & 31u);
}
/* 32 query elements k0..k0+31 of a view as bit planes (bit j = element k0+j) */
__device__ __forceinline__ void lf_q32(const LfDev &d, const LfQView &v, int64_t k0, uint32_t &lo, uint32_t &hi, uint32_t &nn)
{
    if (v.dir > 0) {
        int64_t B = v.bit0 + k0;
        lo = lf_bits32(d.plo, B); hi = lf_bits32(d.phi, B); nn = lf_bits32(d.pnn, B);
    } else {
        int64_t B = v.bit0 - k0 - 31;
        lo = __brev(lf_bits32(d.plo, B)); hi = __brev(lf_bits32(d.phi, B)); nn = __brev(lf_bits32(d.pnn, B));
    }
    lo ^= v.comp; hi ^= v.comp;
}

/* ------------------------------------------------------------------------------------------ */
/* k_copy16: host (pinned, mapped) -> device copy done by a kernel                             */
/* ------------------------------------------------------------------------------------------ */
/* The small uploads of a call (seeds, per-chain metadata, follow-up task lists) must not queue in the copy engine
 * behind the hundreds of MB of reads the other lanes of the call have in flight: a kernel reads the pinned source
 * over PCIe itself.  16-byte units; both pointers 16-byte aligned. */
__global__ void __launch_bounds__(256) k_copy16(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

/* ------------------------------------------------------------------------------------------ */
/* k_readback: a few device words -> pinned host memory, written by a kernel                   */
/* ------------------------------------------------------------------------------------------ */
/* The class counts and totals the host waits for once per batch are ~200 bytes.  As a D2H copy they queue in the copy
 * engine behind whatever else is leaving the device -- the 110 MB of CIGAR / MD text of the early emit while round 3
 * starts cost it 1.5 ms -- so a kernel stores them over PCIe itself (dst is pinned, mapped host memory). */
__global__ void k_readback(const uint32_t *__restrict__ words, uint32_t n_words, const unsigned long long *__restrict__ a, const unsigned long long *__restrict__ b,
                           uint32_t *dst, uint32_t off_a /* in 8-byte units */, uint32_t off_b)
{
    for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = words[i];
    if (threadIdx.x == 0) {
        if (a) ((unsigned long long *)dst)[off_a] = *a;
        if (b) ((unsigned long long *)dst)[off_b] = *b;
    }
#if defined(__CUDA_ARCH__)
    __threadfence_system();
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* k_pack_reads                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* One block per read, 2048 bases per pass: the pass's bytes come in as aligned 16-byte loads (coalesced, whatever the
 * read's own alignment) through shared memory; every thread then turns its 16 bases into 16 bits of each plane with
 * word-wide arithmetic -- the 2-bit code is ((c >> 1) ^ (c >> 2)) & 3 (A 0, C 1, G 2, T 3) for four bytes at a time, a
 * byte is valid if it equals "ACGT"[code] (one PRMT for four bytes), and four flag bits at byte positions 0, 8, 16, 24
 * are gathered into a nibble by one multiplication -- and pairs of threads store whole plane words.
 * History: one base per lane and ballot, 64-bit indices: 2.0 warp instructions per base (0.53 ms for a config-2 chunk,
 * 8 % of the HBM bound); four bases per lane in flight: 0.6 per base (0.29 ms); this version: ~0.2 per base. */
__device__ __forceinline__ uint32_t lf_gather4(uint32_t y)
{ /* bits 0, 8, 16, 24 of y -> bits 0..3 */
    return (y * 0x01020408u) >> 24;
}
__device__ __forceinline__ void lf_pack4(uint32_t x, uint32_t &lo, uint32_t &hi, uint32_t &nn)
{ /* four bases (bytes of x, first base in the low byte) -> four bits of each plane */
    const uint32_t t = (x >> 1) ^ (x >> 2);
    lo = lf_gather4(t & 0x01010101u);
    hi = ((t & 0x02020202u) * 0x00810204u) >> 24;                          /* the same gather for bits 1, 9, 17, 25 */
    const uint32_t u = t & 0x03030303u;                                   /* codes, one per byte */
    const uint32_t v = u | (u >> 4);                                      /* bytes 0 and 2 hold the selector nibble pairs */
    const uint32_t sel = (v & 0xffu) | ((v >> 8) & 0xff00u);              /* PRMT selector: nibble k = code of byte k */
    const uint32_t diff = x ^ __byte_perm(0x54474341u, 0u, sel);          /* zero bytes where the base is upper-case ACGT */
    const uint32_t nz = (diff | ((diff & 0x7f7f7f7fu) + 0x7f7f7f7fu)) >> 7; /* bit 0 of each byte: the byte of diff is not zero */
    nn = lf_gather4(nz & 0x01010101u);
}
__global__ void __launch_bounds__(128) k_pack_reads(LfDev d)
{
    __shared__ uint4 s_tile[130];
    const uint32_t r = blockIdx.x;
    if (r >= d.n_reads) return;
    const uint64_t b0 = d.read_off[r];
    const uint32_t L = (uint32_t)(d.read_off[r + 1] - b0);
    const uint64_t po = lf_plane_word_off(b0, r);
    const uint32_t tid = threadIdx.x, mis = (uint32_t)(b0 & 15ull);
    const uint4 *__restrict__ src = (const uint4 *)(d.bases + (b0 - mis));   /* the allocation is 256-byte aligned and 64 bytes longer than the reads */
    const uint32_t *s_words = (const uint32_t *)s_tile;
    const uint32_t nw = (L + 31u) >> 5;
    for (uint32_t t0 = 0; t0 < L; t0 += 2048u) {
        const uint32_t c0 = t0 >> 4, nchunk = ((L - t0 < 2048u ? L - t0 : 2048u) + mis + 15u) >> 4;   /* aligned chunks that hold bases of this pass */
        if (tid < nchunk) s_tile[tid] = __ldg(src + c0 + tid);
        if (tid == 0 && nchunk > 128u) s_tile[128] = __ldg(src + c0 + 128u);
        __syncthreads();
        const uint32_t i0 = t0 + 16u * tid;                     /* first base of this thread */
        uint32_t lo = 0, hi = 0, nn = 0xffffu;
        if (i0 < L) {
            const uint32_t wi = (mis >> 2) + 4u * tid, sh = 8u * (mis & 3u);
            uint32_t a[5];
#pragma unroll
            for (int j = 0; j < 5; j++) a[j] = s_words[wi + (uint32_t)j];
            const uint32_t left = L - i0;                       /* bases from i0 to the end of the read */
            lo = 0; hi = 0; nn = 0;
            if (left >= 16u) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t l4, h4, n4;
                    lf_pack4(__funnelshift_r(a[j], a[j + 1], sh), l4, h4, n4);
                    lo |= l4 << (4 * j); hi |= h4 << (4 * j); nn |= n4 << (4 * j);
                }
            } else {   /* the last bases of the read: bytes past its end count as byte 0, i.e. code 0 and not a base */
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t x = __funnelshift_r(a[j], a[j + 1], sh);
                    const uint32_t have = left > 4u * (uint32_t)j ? left - 4u * (uint32_t)j : 0u;
                    if (have < 4u) x &= (1u << (8u * have)) - 1u;
                    uint32_t l4, h4, n4;
                    lf_pack4(x, l4, h4, n4);
                    lo |= l4 << (4 * j); hi |= h4 << (4 * j); nn |= n4 << (4 * j);
                }
            }
        }
        /* even threads own a plane word: their 16 bits and the next thread's */
        const uint32_t lo2 = __shfl_down_sync(LF_FULL, lo, 1), hi2 = __shfl_down_sync(LF_FULL, hi, 1), nn2 = __shfl_down_sync(LF_FULL, nn, 1);
        const uint32_t w = (t0 >> 5) + (tid >> 1);
        if (!(tid & 1u) && w < nw) {
            d.plo[po + w] = lo | (lo2 << 16);
            d.phi[po + w] = hi | (hi2 << 16);
            d.pnn[po + w] = nn | (nn2 << 16);
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------ */
/* k_align_prep                                                                               */
/* ------------------------------------------------------------------------------------------ */
__host__ __device__ __forceinline__ unsigned long long lf_large_planes_bytes(uint32_t q, uint32_t t)
{ /* bytes of traceback planes the biggest leaf below (q,t) can need: 8 B per (32-row word x step) */
    unsigned long long n = (q + 31u) / 32u + 8u; /* lanes are padded to whole groups of up to 8 words */
    unsigned long long full = n * ((unsigned long long)t + 32ull) * 8ull;
    unsigned long long cap = (1ull << 20) + n * 512ull + 8192ull;
    return full < cap ? full : cap;
}

__global__ void __launch_bounds__(256) k_align_prep(LfDev d, uint32_t *keys, uint32_t *idx, uint32_t *slot_words, uint32_t *scr_bytes, LfCounters *cnt)
{
    /* per-block partial sums in shared memory, one global atomic per counter per block (2.4 M tasks
     * hitting the same four addresses cost 5.5 ms in the first version of this kernel) */
    __shared__ uint32_t s_hist[LF_NCLS];
    __shared__ uint32_t s_maxq, s_maxt;
    __shared__ unsigned long long s_maxp, s_cells, s_wc, s_swc;
    if (threadIdx.x < LF_NCLS) s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_maxq = 0; s_maxt = 0; s_maxp = 0; s_cells = 0; s_wc = 0; s_swc = 0; }
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    /* this thread's contributions; reduced over the warp below (256 threads hitting the same shared counters with 64-bit
     * atomics made this kernel 10x slower than its 110 MB of traffic) */
    unsigned long long my_cells = 0, my_wc = 0, my_swc = 0, my_maxp = 0;
    uint32_t my_maxq = 0, my_maxt = 0;
    int cls = -1;
    if (i < d.n_tasks) {
        lf_align_task t = d.tasks[i];
        cls = LF_CLS_BAD;
        uint32_t scr = 0, slot = 0;
        bool ok = t.q_len >= 1 && t.t_len >= 1 && t.read_id < d.n_reads && t.mode <= LF_MODE_SHW;
        if (ok) {
            uint64_t L = d.read_off[t.read_id + 1] - d.read_off[t.read_id];
            ok = (uint64_t)t.q_off + t.q_len <= L && (int64_t)t.t_off + (int64_t)t.t_len <= d.l_pac
                 && (uint64_t)t.q_len + t.t_len < (1ull << 31);
        }
        if (ok) {
            uint32_t nwords = (t.q_len + 31u) >> 5;
            int sc = lf_small_class(nwords);
            const bool leaf = lf_is_leaf(t.q_len, t.t_len), nopath = (t.flags & LF_F_NO_PATH) != 0;
            if (sc >= 0 && leaf) {
                cls = 2 * sc + (t.mode == LF_MODE_SHW ? 1 : 0);
                const bool br = sc >= 4 && (d.bandreg >> (sc - 4) & 1u) && t.mode == LF_MODE_NW && lf_bandreg_eligible(t.q_len, t.t_len, sc);
                if (br) cls = LF_CLS_BANDREG0 + sc - 4;
                scr = lf_k1_ckpt_bytes(t.t_len, lf_small_nw(sc));
                my_swc = (unsigned long long)nwords * t.t_len;
                /* 257 .. 512 rows off the diagonal or in prefix mode: a few long tasks, one thread each would be the tail of the step */
                if (!br && !nopath && sc >= 6 && (d.groupk & 4u)) { cls = LF_CLS_GP16; scr = 0; }
            } else if (nopath && (d.groupk & 1u) && nwords > 16u && nwords <= 256u) {
                cls = nwords <= 32u ? LF_CLS_GD32 : nwords <= 64u ? LF_CLS_GD64 : nwords <= 128u ? LF_CLS_GD128 : LF_CLS_GD256;
            } else if (!nopath && leaf && (d.groupk & 2u) && nwords > 16u && nwords <= 64u) {
                cls = nwords <= 32u ? LF_CLS_GP32 : LF_CLS_GP64;
            } else {
                cls = LF_CLS_LARGE;
                my_maxq = t.q_len; my_maxt = t.t_len; my_maxp = lf_large_planes_bytes(t.q_len, t.t_len);
            }
            if (cls >= LF_CLS_GP16 && cls <= LF_CLS_GP64) atomicMax(&cnt->gmax_t[cls - LF_CLS_GP16], t.t_len);   /* a few thousand tasks per chunk */
            slot = nopath ? 0u : (t.q_len + t.t_len + 15u) >> 4;
            my_cells = (unsigned long long)t.q_len * t.t_len;
            my_wc = (unsigned long long)nwords * t.t_len;
        } else {
            lf_align_result r; r.edit_distance = -1; r.end_location = -1; r.ops_off = 0; r.ops_len = 0; r.status = LF_ERR_BAD_ARG;
            d.res[i] = r;
        }
        /* class-major, long targets first; the target length is quantised to 1/8 octave so that a warp's
         * 32 tasks run nearly the same number of columns, and inside a bucket the (stable) sort keeps the
         * submission order, i.e. neighbouring lanes work on neighbouring reads and result slots */
        const uint32_t tl = t.t_len ? t.t_len : 1u;
        const uint32_t e = 31u - (uint32_t)__clz((int)tl);
        constexpr uint32_t MB = LF_BUCKET_BITS, MM = (1u << MB) - 1u;
        const uint32_t m = e >= MB ? (tl >> (e - MB)) & MM : (tl << (MB - e)) & MM;
        keys[i] = ((uint32_t)cls << 27) | ((((32u << MB) - 1u) - ((e << MB) + m)) << LF_KEY_SHIFT);
        idx[i] = i;
        slot_words[i] = slot;
        scr_bytes[i] = scr;
    }
    {   /* warp reduction, then one shared-memory atomic per warp and counter */
        const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            my_cells += __shfl_down_sync(LF_FULL, my_cells, o);
            my_wc += __shfl_down_sync(LF_FULL, my_wc, o);
            my_swc += __shfl_down_sync(LF_FULL, my_swc, o);
            const unsigned long long p = __shfl_down_sync(LF_FULL, my_maxp, o);
            const uint32_t mq = __shfl_down_sync(LF_FULL, my_maxq, o), mt = __shfl_down_sync(LF_FULL, my_maxt, o);
            my_maxp = p > my_maxp ? p : my_maxp; my_maxq = mq > my_maxq ? mq : my_maxq; my_maxt = mt > my_maxt ? mt : my_maxt;
        }
        if (lane == 0) {
            if (my_cells) atomicAdd(&s_cells, my_cells);
            if (my_wc) atomicAdd(&s_wc, my_wc);
            if (my_swc) atomicAdd(&s_swc, my_swc);
            if (my_maxq) atomicMax(&s_maxq, my_maxq);
            if (my_maxt) atomicMax(&s_maxt, my_maxt);
            if (my_maxp) atomicMax(&s_maxp, my_maxp);
        }
#if defined(__CUDA_ARCH__)
        const unsigned same = __match_any_sync(LF_FULL, cls);   /* the lanes of this warp that are in the same class */
        if (cls >= 0 && lane == (uint32_t)(__ffs((int)same) - 1)) atomicAdd(&s_hist[cls], (uint32_t)__popc(same));
#else
        if (cls >= 0) atomicAdd(&s_hist[cls], 1u);
#endif
    }
    __syncthreads();
    if (threadIdx.x < LF_NCLS && s_hist[threadIdx.x]) atomicAdd(&cnt->hist[threadIdx.x], s_hist[threadIdx.x]);
    if (threadIdx.x == 0) {
        if (s_maxq) atomicMax(&cnt->max_q, s_maxq);
        if (s_maxt) atomicMax(&cnt->max_t, s_maxt);
        if (s_maxp) atomicMax(&cnt->max_planes, s_maxp);
        atomicAdd(&cnt->cells, s_cells);
        atomicAdd(&cnt->word_columns, s_wc);
        atomicAdd(&cnt->small_word_columns, s_swc);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* op stream of the thread-per-task kernels                                                    */
/* ------------------------------------------------------------------------------------------ */
/* 2-bit ops, 16 per word, written from the end of the task's slot towards its start (the walk finds the last op of the
 * alignment first).  The accumulator takes each op at its low end (one LEA); every 16th op the word is complete with the
 * first of them on top.  Word indices are 32-bit: the host refuses batches of 2^32 op words. */
struct LfOpSink {
    uint32_t acc, nops, widx;
    __device__ __forceinline__ void init(uint64_t slot_hi) { acc = 0; nops = 0; widx = (uint32_t)(slot_hi >> 4); }
    __device__ __forceinline__ void emit(uint32_t *__restrict__ ops, uint32_t op)
    {
        acc = (acc << 2) | op;
        if ((++nops & 15u) == 0u) ops[--widx] = acc;
    }
    __device__ __forceinline__ void finish(uint32_t *__restrict__ ops) { if (nops & 15u) ops[widx - 1] = acc << (32u - 2u * (nops & 15u)); }
};

/* Walk inside one word-row of a block's window in shared memory (planes [column][window word][2][STRIDE threads], CS words
 * per column): one op per iteration, Up > Left > Diagonal, until the walk leaves the word-row or the block.  cell0 is the
 * index of column c0's plane word for this thread and word-row; the column is carried by the cell index alone.
 * (Counting the matches of the next three diagonal cells first and emitting them together -- 3.2 ops per iteration for 80
 * instructions -- was 3.5 % slower on the config-2 step: lanes leave the loop after different numbers of iterations.) */
template <int STRIDE, int CS>
__device__ __forceinline__ void lf_walk_row(const uint32_t *smem, int cell0, int wrow, int c0, int &i, int &j, LfOpSink &sink, uint32_t *__restrict__ ops)
{
    int cell = cell0 + (j - 1 - c0) * CS;
    int b = (i - 1) & 31;
    do {
        const uint32_t x0 = smem[cell] >> b, x1 = smem[cell + STRIDE] >> b;
        const uint32_t op = (x0 & 1u) | ((x1 & 1u) << 1);
        const int stay_col = (int)(x0 & ~x1 & 1u);  /* op 1: up    */
        const int stay_row = (int)(x1 & ~x0 & 1u);  /* op 2: left  */
        sink.emit(ops, op);
        b = b + stay_row - 1;
        cell = cell + stay_col * CS - CS;
    } while (b >= 0 && cell >= cell0);
    j = c0 + (cell - cell0 + CS) / CS;   /* cell0 - CS (one column left of the block) gives c0 */
    i = wrow * 32 + b + 1;
}

/* ------------------------------------------------------------------------------------------ */
/* k_myers_small: thread-per-task, NW words of 32 rows in registers                            */
/* ------------------------------------------------------------------------------------------ */
template <int NW>
__device__ __forceinline__ void lf_add_chain(const uint32_t (&a)[NW], const uint32_t (&b)[NW], uint32_t (&s)[NW])
{ /* s = a + b over a 32*NW-bit word (portable form; the device build uses the asm chains below) */
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        unsigned long long x = (unsigned long long)a[w] + b[w] + c;
        s[w] = (uint32_t)x;
        c = (uint32_t)(x >> 32);
    }
}
template <int NW>
__device__ __forceinline__ void lf_add_chain_cin(const uint32_t (&a)[NW], const uint32_t (&b)[NW], uint32_t (&s)[NW], uint32_t cin)
{ /* s = a + b + cin (cin in {0,1}) */
    uint32_t c = cin;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        unsigned long long x = (unsigned long long)a[w] + b[w] + c;
        s[w] = (uint32_t)x;
        c = (uint32_t)(x >> 32);
    }
}
/* generated by the snippet in DESIGN.md (carry chains as one asm block each: IADD3 + IADD3.X per word) */
#if defined(__CUDA_ARCH__)
template <> __device__ __forceinline__ void lf_add_chain_cin<1>(const uint32_t (&a)[1], const uint32_t (&b)[1], uint32_t (&s)[1], uint32_t cin) { s[0] = a[0] + b[0] + cin; }
template <> __device__ __forceinline__ void lf_add_chain<2>(const uint32_t (&a)[2], const uint32_t (&b)[2], uint32_t (&s)[2])
{
    asm("add.cc.u32 %0, %2, %4;\n\t"
        "addc.u32 %1, %3, %5;"
        : "=&r"(s[0]),"=&r"(s[1])
        : "r"(a[0]),"r"(a[1]),"r"(b[0]),"r"(b[1]));
}
template <> __device__ __forceinline__ void lf_add_chain<3>(const uint32_t (&a)[3], const uint32_t (&b)[3], uint32_t (&s)[3])
{
    asm("add.cc.u32 %0, %3, %6;\n\t"
        "addc.cc.u32 %1, %4, %7;\n\t"
        "addc.u32 %2, %5, %8;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2])
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(b[0]),"r"(b[1]),"r"(b[2]));
}
template <> __device__ __forceinline__ void lf_add_chain<4>(const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t (&s)[4])
{
    asm("add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, %9;\n\t"
        "addc.cc.u32 %2, %6, %10;\n\t"
        "addc.u32 %3, %7, %11;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3])
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]));
}
template <> __device__ __forceinline__ void lf_add_chain<5>(const uint32_t (&a)[5], const uint32_t (&b)[5], uint32_t (&s)[5])
{
    asm("add.cc.u32 %0, %5, %10;\n\t"
        "addc.cc.u32 %1, %6, %11;\n\t"
        "addc.cc.u32 %2, %7, %12;\n\t"
        "addc.cc.u32 %3, %8, %13;\n\t"
        "addc.u32 %4, %9, %14;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3]),"=&r"(s[4])
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]),"r"(b[4]));
}
template <> __device__ __forceinline__ void lf_add_chain<6>(const uint32_t (&a)[6], const uint32_t (&b)[6], uint32_t (&s)[6])
{
    asm("add.cc.u32 %0, %6, %12;\n\t"
        "addc.cc.u32 %1, %7, %13;\n\t"
        "addc.cc.u32 %2, %8, %14;\n\t"
        "addc.cc.u32 %3, %9, %15;\n\t"
        "addc.cc.u32 %4, %10, %16;\n\t"
        "addc.u32 %5, %11, %17;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3]),"=&r"(s[4]),"=&r"(s[5])
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(a[5]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]),"r"(b[4]),"r"(b[5]));
}
template <> __device__ __forceinline__ void lf_add_chain<8>(const uint32_t (&a)[8], const uint32_t (&b)[8], uint32_t (&s)[8])
{
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3]),"=&r"(s[4]),"=&r"(s[5]),"=&r"(s[6]),"=&r"(s[7])
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(a[5]),"r"(a[6]),"r"(a[7]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]),"r"(b[4]),"r"(b[5]),"r"(b[6]),"r"(b[7]));
}
template <> __device__ __forceinline__ void lf_add_chain<12>(const uint32_t (&a)[12], const uint32_t (&b)[12], uint32_t (&s)[12])
{
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, %33;\n\t"
        "addc.cc.u32 %10, %22, %34;\n\t"
        "addc.u32 %11, %23, %35;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3]),"=&r"(s[4]),"=&r"(s[5]),"=&r"(s[6]),"=&r"(s[7]),"=&r"(s[8]),"=&r"(s[9]),"=&r"(s[10]),"=&r"(s[11])
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(a[5]),"r"(a[6]),"r"(a[7]),"r"(a[8]),"r"(a[9]),"r"(a[10]),"r"(a[11]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]),"r"(b[4]),"r"(b[5]),"r"(b[6]),"r"(b[7]),"r"(b[8]),"r"(b[9]),"r"(b[10]),"r"(b[11]));
}
template <> __device__ __forceinline__ void lf_add_chain<16>(const uint32_t (&a)[16], const uint32_t (&b)[16], uint32_t (&s)[16])
{
    asm("add.cc.u32 %0, %16, %32;\n\t"
        "addc.cc.u32 %1, %17, %33;\n\t"
        "addc.cc.u32 %2, %18, %34;\n\t"
        "addc.cc.u32 %3, %19, %35;\n\t"
        "addc.cc.u32 %4, %20, %36;\n\t"
        "addc.cc.u32 %5, %21, %37;\n\t"
        "addc.cc.u32 %6, %22, %38;\n\t"
        "addc.cc.u32 %7, %23, %39;\n\t"
        "addc.cc.u32 %8, %24, %40;\n\t"
        "addc.cc.u32 %9, %25, %41;\n\t"
        "addc.cc.u32 %10, %26, %42;\n\t"
        "addc.cc.u32 %11, %27, %43;\n\t"
        "addc.cc.u32 %12, %28, %44;\n\t"
        "addc.cc.u32 %13, %29, %45;\n\t"
        "addc.cc.u32 %14, %30, %46;\n\t"
        "addc.u32 %15, %31, %47;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3]),"=&r"(s[4]),"=&r"(s[5]),"=&r"(s[6]),"=&r"(s[7]),"=&r"(s[8]),"=&r"(s[9]),"=&r"(s[10]),"=&r"(s[11]),"=&r"(s[12]),"=&r"(s[13]),"=&r"(s[14]),"=&r"(s[15])
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(a[5]),"r"(a[6]),"r"(a[7]),"r"(a[8]),"r"(a[9]),"r"(a[10]),"r"(a[11]),"r"(a[12]),"r"(a[13]),"r"(a[14]),"r"(a[15]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]),"r"(b[4]),"r"(b[5]),"r"(b[6]),"r"(b[7]),"r"(b[8]),"r"(b[9]),"r"(b[10]),"r"(b[11]),"r"(b[12]),"r"(b[13]),"r"(b[14]),"r"(b[15]));
}
template <> __device__ __forceinline__ void lf_add_chain_cin<2>(const uint32_t (&a)[2], const uint32_t (&b)[2], uint32_t (&s)[2], uint32_t cin)
{
    uint32_t tmp;
    asm("add.cc.u32 %2, %7, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %3, %5;\n\t"
        "addc.u32 %1, %4, %6;"
        : "=&r"(s[0]),"=&r"(s[1]), "=&r"(tmp)
        : "r"(a[0]),"r"(a[1]),"r"(b[0]),"r"(b[1]), "r"(cin));
}
template <> __device__ __forceinline__ void lf_add_chain_cin<4>(const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t (&s)[4], uint32_t cin)
{
    uint32_t tmp;
    asm("add.cc.u32 %4, %13, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.u32 %3, %8, %12;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3]), "=&r"(tmp)
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]), "r"(cin));
}
template <> __device__ __forceinline__ void lf_add_chain_cin<8>(const uint32_t (&a)[8], const uint32_t (&b)[8], uint32_t (&s)[8], uint32_t cin)
{
    uint32_t tmp;
    asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.u32 %7, %16, %24;"
        : "=&r"(s[0]),"=&r"(s[1]),"=&r"(s[2]),"=&r"(s[3]),"=&r"(s[4]),"=&r"(s[5]),"=&r"(s[6]),"=&r"(s[7]), "=&r"(tmp)
        : "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(a[5]),"r"(a[6]),"r"(a[7]),"r"(b[0]),"r"(b[1]),"r"(b[2]),"r"(b[3]),"r"(b[4]),"r"(b[5]),"r"(b[6]),"r"(b[7]), "r"(cin));
}
#endif

/* max of v over the lanes of the warp that execute this together (any superset value is valid for the callers: they
 * pick a code variant that covers at least what every lane needs) */
__device__ __forceinline__ int lf_converged_max(int v)
{
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(__activemask(), v);
#else
    return v;
#endif
}

__device__ __forceinline__ int lf_warp_max(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) { const int t = __shfl_xor_sync(LF_FULL, v, o); v = t > v ? t : v; }
    return v;
}

/* Advance the whole column by one target symbol (Myers 1999 / Hyyro 2003 recurrences on one long
 * word; the per-block hin/hout of edlib's calculateBlock, edlib.cpp:335-370, become the add carry
 * and the bits funnel-shifted between words): 12 integer instructions per word.
 * STORE additionally writes, for the WIN words starting at wtop, the two traceback planes of this
 * column: op = 1 (up) if Pv', else 2 (left) if Ph, else 0/3 by Eq -- plane0 = low op bit, plane1 =
 * high op bit.  The recompute instantiates this with NW = the number of words down to the window
 * (words below it can never be reached by the walk) and only the last SPAN words can be window words. */
template <int NW, bool SHW, bool STORE, int WIN, int SPAN>
__device__ __forceinline__ void lf_k1_column(uint32_t (&Pv)[NW], uint32_t (&Mv)[NW], const uint32_t (&qlo)[NW],
                                             const uint32_t (&qhi)[NW], const uint32_t (&qnn)[NW], uint32_t slo, uint32_t shi, int &score,
                                             int wl, uint32_t bl, uint32_t *sm, int wtop)
{
    uint32_t Eq[NW], a[NW], sum[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) {
        Eq[w] = ~lf_neq(qlo[w], qhi[w], qnn[w], slo, shi);   /* two LOP3s; the negation folds into the users */
        a[w] = Eq[w] & Pv[w];
    }
    lf_add_chain<NW>(a, Pv, sum);
    uint32_t pPh = 0x80000000u, pMh = 0u; /* row 0 of a global alignment grows by one per column */
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const uint32_t Xh = (sum[w] ^ Pv[w]) | Eq[w];
        const uint32_t Ph = Mv[w] | ~(Xh | Pv[w]);
        const uint32_t Mh = Pv[w] & Xh;
        const uint32_t Xv = Eq[w] | Mv[w];
        if (SHW) { if (w == wl) score += (int)((Ph >> bl) & 1u) - (int)((Mh >> bl) & 1u); }
        const uint32_t Phs = __funnelshift_l(pPh, Ph, 1), Mhs = __funnelshift_l(pMh, Mh, 1);
        pPh = Ph; pMh = Mh;
        const uint32_t nPv = Mhs | ~(Xv | Phs);
        const uint32_t nMv = Phs & Xv;
        if (STORE && w + SPAN >= NW) {
            const unsigned wi = (unsigned)(w - wtop);
            if (wi < (unsigned)WIN) {
                const uint32_t diagx = ~(nPv | Ph | Eq[w]);              /* diagonal step over a mismatch */
                sm[(wi * 2 + 0) * LF_K1_BLOCK] = nPv | diagx;         /* ops 1, 3 */
                sm[(wi * 2 + 1) * LF_K1_BLOCK] = (~nPv & Ph) | diagx; /* ops 2, 3 */
            }
        }
        Pv[w] = nPv; Mv[w] = nMv;
    }
}

/* Recompute columns [c0, c1) from the state in Pv/Mv, touching only the first NWC words. */
template <int NW, int NWC, int WIN, int SPAN>
__device__ __forceinline__ void lf_k1_recompute(uint32_t (&Pv)[NW], uint32_t (&Mv)[NW], const uint32_t (&qlo)[NW], const uint32_t (&qhi)[NW],
                                                const uint32_t (&qnn)[NW], LfTStreamCol &tc, int ncols, uint32_t *smt, int wtop)
{
    static_assert(NWC <= NW, "");
    uint32_t (&P)[NWC] = reinterpret_cast<uint32_t (&)[NWC]>(Pv);
    uint32_t (&M)[NWC] = reinterpret_cast<uint32_t (&)[NWC]>(Mv);
    const uint32_t (&L)[NWC] = reinterpret_cast<const uint32_t (&)[NWC]>(qlo);
    const uint32_t (&H)[NWC] = reinterpret_cast<const uint32_t (&)[NWC]>(qhi);
    const uint32_t (&N)[NWC] = reinterpret_cast<const uint32_t (&)[NWC]>(qnn);
    int dummy = 0;
    for (int c = 0; c < ncols; c++) {
        uint32_t slo, shi;
        tc.next_masks(slo, shi);
        lf_k1_column<NWC, false, true, WIN, SPAN>(P, M, L, H, N, slo, shi, dummy, 0, 0u, smt + (size_t)c * (WIN * 2 * LF_K1_BLOCK), wtop);
    }
}

template <int NW, bool SHW>
__global__ void __launch_bounds__(LF_K1_BLOCK) k_myers_small(LfDev d, const uint32_t *__restrict__ order, uint32_t first, uint32_t count, const uint32_t *__restrict__ retry_count)
{
    constexpr int WIN = NW < 2 ? 1 : 2;
    constexpr int C = LF_K1_C;
    LF_DYN_SMEM(uint32_t, smem); /* [C][WIN][2][LF_K1_BLOCK] */
    const uint32_t tid = threadIdx.x;
    const uint32_t gi = blockIdx.x * LF_K1_BLOCK + tid;
    if (gi >= count) return;
    if (retry_count && gi >= *retry_count) return; /* `order` is then the dense list k_myers_band appended its uncertified tasks to */
    const uint32_t ti = order[first + gi];
    const lf_align_task task = d.tasks[ti];
    const int q = (int)task.q_len, t = (int)task.t_len;
    LfQView qv; LfTView tv;
    lf_task_views(d, task, qv, tv);

    uint32_t qlo[NW], qhi[NW], qnn[NW], Pv[NW], Mv[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) {
        if (w * 32 < q) lf_q32(d, qv, (int64_t)w * 32, qlo[w], qhi[w], qnn[w]);
        else { qlo[w] = 0; qhi[w] = 0; qnn[w] = 0xffffffffu; }
        Pv[w] = 0xffffffffu; Mv[w] = 0u;
    }
    const int wl = (q - 1) >> 5;
    const uint32_t bl = (uint32_t)(q - 1) & 31u;
    int score = q, best = q, bestc = -1;
    uint2 *ck = (uint2 *)(d.scratch + d.scr_off[ti]);

    /* ---- forward pass: distance (+ first best prefix for SHW), checkpoints every C columns ---- */
    LfTStreamCol tc;
    tc.init(d.pac, tv.t0, tv.dir);
    for (int c = 0; c < t; c++) {
        if ((c & (C - 1)) == 0 && c) {
            uint2 *dst = ck + (size_t)(c / C - 1) * NW;
#pragma unroll
            for (int w = 0; w < NW; w++) dst[w] = make_uint2(Pv[w], Mv[w]);
        }
        uint32_t slo, shi;
        tc.next_masks(slo, shi);
        lf_k1_column<NW, SHW, false, WIN, 0>(Pv, Mv, qlo, qhi, qnn, slo, shi, score, wl, bl, nullptr, 0);
        if (SHW) { if (score < best) { best = score; bestc = c; } }
    }
    int ed, end;
    if (SHW) { ed = best; end = bestc; }
    else {
        ed = t;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            uint32_t m = w < wl ? 0xffffffffu : w == wl ? (0xffffffffu >> (31u - bl)) : 0u;
            ed += __popc(Pv[w] & m) - __popc(Mv[w] & m);
        }
        end = t - 1;
    }
    lf_align_result r;
    r.edit_distance = ed; r.end_location = end; r.status = 0;
    const uint64_t slot_hi = d.slot_end[ti] * 16ull; /* one past the last op position of the slot */
    if (task.flags & LF_F_NO_PATH) { r.ops_off = slot_hi; r.ops_len = 0; d.res[ti] = r; return; }

    /* ---- traceback: recompute 16-column blocks from their checkpoint, keep a WIN-word window of
     *      the op planes in shared memory, walk Up > Left > Diagonal (edlib.cpp:950, :984, :1015) ---- */
    LfOpSink sink;
    sink.init(slot_hi);
    int i = q, j = end + 1;
    uint32_t *smt = smem + tid;
    constexpr int CS = WIN * 2 * LF_K1_BLOCK; /* shared-memory words per column */
    while (i > 0 && j > 0) {
        const int c1 = j, c0 = ((j - 1) / C) * C;
        const int whi = (i - 1) >> 5;
        const int wtop = whi - WIN + 1 > 0 ? whi - WIN + 1 : 0;
        if (c0 == 0) {
#pragma unroll
            for (int w = 0; w < NW; w++) { Pv[w] = 0xffffffffu; Mv[w] = 0u; }
        } else {
            const uint2 *src = ck + (size_t)(c0 / C - 1) * NW;
#pragma unroll
            for (int w = 0; w < NW; w++) { if (w <= whi) { uint2 v = src[w]; Pv[w] = v.x; Mv[w] = v.y; } }
        }
        tc.init(d.pac, tv.t0 + (int64_t)tv.dir * c0, tv.dir);
        {
            /* words 0..whi only, in compile-time sized variants (a quarter of the class width each) */
            constexpr int G = NW >= 8 ? NW / 4 : NW == 6 ? 2 : 1;
            constexpr int SPAN = NW;   /* every computed word may hold the window: the variant is the warp's, not the smallest that covers this lane's row */
            const int ncols = c1 - c0;
            /* one variant for the lanes that are here together (the deepest row among them decides): lanes choosing
             * different variants would run them one after the other */
            const int ws = lf_converged_max(whi);
            if (NW > G && ws < G) lf_k1_recompute<NW, G, WIN, SPAN>(Pv, Mv, qlo, qhi, qnn, tc, ncols, smt, wtop);
            else if (NW > 2 * G && ws < 2 * G) lf_k1_recompute<NW, (2 * G < NW ? 2 * G : NW), WIN, SPAN>(Pv, Mv, qlo, qhi, qnn, tc, ncols, smt, wtop);
            else if (NW > 3 * G && ws < 3 * G) lf_k1_recompute<NW, (3 * G < NW ? 3 * G : NW), WIN, SPAN>(Pv, Mv, qlo, qhi, qnn, tc, ncols, smt, wtop);
            else lf_k1_recompute<NW, NW, WIN, SPAN>(Pv, Mv, qlo, qhi, qnn, tc, ncols, smt, wtop);
        }
        /* walk inside the window, one word-row at a time */
        const int rowmin = wtop * 32;
        while (i > 0 && j > c0 && (i - 1) >= rowmin) {
            const int wrow = (i - 1) >> 5;
            lf_walk_row<LF_K1_BLOCK, CS>(smem, (int)tid + (wrow - wtop) * 2 * LF_K1_BLOCK, wrow, c0, i, j, sink, d.ops);
        }
    }
    while (i > 0) { sink.emit(d.ops, 1u); i--; } /* left column: the rest of the query is inserted   */
    while (j > 0) { sink.emit(d.ops, 2u); j--; } /* top row: the rest of the target is deleted        */
    sink.finish(d.ops);
    const uint32_t nops = sink.nops;
    const uint64_t p = slot_hi - nops;
    r.ops_off = p; r.ops_len = (uint32_t)(slot_hi - p);
    d.res[ti] = r;
}

/* ------------------------------------------------------------------------------------------ */
/* k_myers_large: warp-per-task wavefront + Hirschberg                                         */
/* ------------------------------------------------------------------------------------------ */
struct LfLargeCfg {
    uint8_t *base;              /* scratch of all warp slots */
    unsigned long long stride;  /* bytes per warp slot */
    unsigned long long off_hb, off_L, off_R, off_opsb, off_stack; /* planes start at 0 */
    uint32_t *queue;            /* work counter */
};

enum { LF_PASS_STORE = 1, LF_PASS_SHW = 2, LF_PASS_COL = 4 };

struct LfPassOut { int ed, best, bestc; };

/* words of 32 rows each lane owns in a wavefront pass over a query of ql rows */
__host__ __device__ __forceinline__ int lf_wpl(int ql)
{
    const int n = (ql + 31) >> 5;
    return n <= 32 ? 1 : n <= 64 ? 2 : n <= 128 ? 4 : 8;
}

#define LF_WAVE_CB 8 /* columns a lane advances per wavefront step */

/* One wavefront pass over (query view, target view).  Lane l of strip s owns WPL consecutive words
 * (32*WPL rows) as one long word (carry chain inside the lane) and at step k works on the block of
 * LF_WAVE_CB columns number k-l, so the __shfl_up that hands the 8 hout values (2 bits each) to lane
 * l+1 is paid once per 8 columns and the critical path is ~(t + 8*lanes) column latencies instead of
 * t shuffle round trips.  hout enters the next lane as the add's carry-in and the bits shifted into
 * Ph/Mh.  Every lane reads the target through its own cursor.  Strips of 32*WPL words run one after
 * the other, chained through hb[] (only queries above 8192 rows need a second strip).
 * Returns D(ql, tl) in .ed; with LF_PASS_SHW also the minimum of the last row and the first column
 * reaching it; LF_PASS_COL writes D(x, tl), x = 0..ql, to col[]; LF_PASS_STORE writes the traceback
 * planes as uint2 at planes[strip_base + (column*nv + lane)*WPL + k]. */
template <int WPL>
__device__ __forceinline__ LfPassOut lf_wave_pass_t(const LfDev &d, const LfQView &qv, int ql, const LfTView &tv, int tl, int flags,
                                                    uint2 *planes, int8_t *hb, int32_t *col)
{
    constexpr int CB = LF_WAVE_CB;
    const int lane = threadIdx.x & 31;
    const int n = (ql + 31) >> 5;
    const int SW = 32 * WPL;                   /* words per strip */
    const int S = (n + SW - 1) / SW;
    const int wl = (ql - 1) >> 5;
    const uint32_t bl = (uint32_t)(ql - 1) & 31u;
    int score = ql, best = ql, bestc = -1;     /* tracked by the lane owning row ql-1 */
    int colbase = tl;                          /* D(first row of the strip, tl) carried across strips */
    unsigned long long sbase = 0;
    if ((flags & LF_PASS_COL) && lane == 0) col[0] = tl;
    const int nblocks = (tl + CB - 1) / CB;
    for (int s = 0; s < S; s++) {
        const int w0 = (s * 32 + lane) * WPL;
        const int nvw = n - s * SW < SW ? n - s * SW : SW;
        const int nv = (nvw + WPL - 1) / WPL;
        const bool valid = lane < nv;
        uint32_t lo[WPL], hi[WPL], nn[WPL], Pv[WPL], Mv[WPL];
#pragma unroll
        for (int k = 0; k < WPL; k++) {
            lo[k] = 0; hi[k] = 0; nn[k] = 0xffffffffu;
            if (valid && w0 + k < n) lf_q32(d, qv, (int64_t)(w0 + k) * 32, lo[k], hi[k], nn[k]);
            Pv[k] = 0xffffffffu; Mv[k] = 0u;
        }
        uint32_t pay = 0; /* (hout+1) of the 8 columns this lane did in the previous step, 2 bits each */
        const int nsteps = nblocks + nv - 1;
        LfTCursor tc;
        tc.init(d.pac, tv.t0, tv.dir); /* every lane starts at column 0 when its first block arrives */
        for (int step = 0; step < nsteps; step++) {
            uint32_t in = __shfl_up_sync(LF_FULL, pay, 1);
            const int cb = step - lane;
            const bool act = valid && cb >= 0 && cb < nblocks;
            pay = 0;
            if (act) {
                const int cbase = cb * CB;
                if (lane == 0) { /* row 0 of the strip: +1 per column, or the previous strip's bottom row */
                    in = 0xaaaau;
                    if (s > 0) { in = 0; for (int ci = 0; ci < CB && cbase + ci < tl; ci++) in |= (uint32_t)((int)hb[cbase + ci] + 1) << (2 * ci); }
                }
                const int ncol = tl - cbase < CB ? tl - cbase : CB;
                for (int ci = 0; ci < ncol; ci++) {
                    const int c = cbase + ci;
                    const int hin = (int)((in >> (2 * ci)) & 3u) - 1;
                    uint32_t slo, shi;
                    tc.next_masks(slo, shi);
                    uint32_t Eq[WPL], a[WPL], sum[WPL];
#pragma unroll
                    for (int k = 0; k < WPL; k++) {
                        Eq[k] = ~((lo[k] ^ slo) | (hi[k] ^ shi) | nn[k]);
                        a[k] = Eq[k] & Pv[k];
                    }
                    const uint32_t hneg = hin < 0 ? 1u : 0u;
                    lf_add_chain_cin<WPL>(a, Pv, sum, hneg);   /* hin = -1 enters as the carry-in */
                    uint32_t pPh = hin > 0 ? 0x80000000u : 0u, pMh = hneg << 31;
                    uint2 *dst = planes + sbase + ((unsigned long long)c * nv + lane) * WPL;
#pragma unroll
                    for (int k = 0; k < WPL; k++) {
                        const uint32_t Xh = (sum[k] ^ Pv[k]) | Eq[k] | (k == 0 ? hneg : 0u);
                        const uint32_t Ph = Mv[k] | ~(Xh | Pv[k]);
                        const uint32_t Mh = Pv[k] & Xh;
                        const uint32_t Xv = Eq[k] | Mv[k];
                        if ((flags & LF_PASS_SHW) && w0 + k == wl) {
                            score += (int)((Ph >> bl) & 1u) - (int)((Mh >> bl) & 1u);
                            if (score < best) { best = score; bestc = c; }
                        }
                        const uint32_t Phs = __funnelshift_l(pPh, Ph, 1), Mhs = __funnelshift_l(pMh, Mh, 1);
                        pPh = Ph; pMh = Mh;
                        const uint32_t nPv = Mhs | ~(Xv | Phs);
                        const uint32_t nMv = Phs & Xv;
                        if (flags & LF_PASS_STORE) {
                            const uint32_t diagx = ~(nPv | Ph | Eq[k]);
                            dst[k] = make_uint2(nPv | diagx, (~nPv & Ph) | diagx);
                        }
                        Pv[k] = nPv; Mv[k] = nMv;
                    }
                    const int hout = (int)(pPh >> 31) - (int)(pMh >> 31);
                    pay |= (uint32_t)(hout + 1) << (2 * ci);
                    if (lane == nv - 1 && s + 1 < S) hb[c] = (int8_t)hout;
                }
            }
        }
        sbase += (unsigned long long)tl * nv * WPL;
        /* last column of this strip: vertical deltas -> absolute values */
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < WPL; k++) {
            const int w = w0 + k;
            const uint32_t m = !valid ? 0u : w < wl ? 0xffffffffu : w == wl ? (0xffffffffu >> (31u - bl)) : 0u;
            cnt += __popc(Pv[k] & m) - __popc(Mv[k] & m);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(LF_FULL, incl, o); if (lane >= o) incl += v; }
        if (flags & LF_PASS_COL) {
            if (valid) {
                int v = colbase + incl - cnt;
#pragma unroll
                for (int k = 0; k < WPL; k++) {
                    const int w = w0 + k;
                    const int rows = ql - w * 32 < 32 ? ql - w * 32 : 32;
                    for (int b = 0; b < rows; b++) { v += (int)((Pv[k] >> b) & 1u) - (int)((Mv[k] >> b) & 1u); col[w * 32 + b + 1] = v; }
                }
            }
        }
        colbase += __shfl_sync(LF_FULL, incl, 31);
        __syncwarp();
    }
    LfPassOut o;
    o.ed = colbase;
    const int owner = ((wl % SW) / WPL) & 31;
    o.best = __shfl_sync(LF_FULL, best, owner);
    o.bestc = __shfl_sync(LF_FULL, bestc, owner);
    return o;
}

/* ------------------------------------------------------------------------------------------ */
/* lf_gwave: the wavefront pass for a GROUP of LANES lanes (4, 8, 16 or 32), flags at compile time   */
/* ------------------------------------------------------------------------------------------ */
/* Same recurrences and lane skew as lf_wave_pass_t, rebuilt for instruction count: the pass variant (store planes /
 * prefix-mode score / last column) is a template argument instead of a run-time test per word; every lane reads the
 * target as a stream (16 symbols per funnel shift); hout travels as two bits (Ph, Mh of the lane's bottom row) per
 * column; Eq is two LOP3s.  A warp holds 32 / LANES groups, each on its own (query, target): all 32 lanes call this
 * together, the step count is the warp's maximum, ql == 0 marks an idle group.  One strip only: the query must fit
 * LANES * WPL words.  Per word-column: 14 instructions of recurrence + 4 when the planes are stored + 2 for the
 * prefix-mode score, and ~12 per lane and column of bookkeeping, amortised over WPL words.
 * Planes layout (as lf_wave_pass_t with one strip): uint2 at planes[(column * nv + lane) * WPL + k]. */
template <int LANES, int WPL, int FLAGS, bool MULTI = false>
__device__ __forceinline__ LfPassOut lf_gwave(const LfDev &d, const LfQView &qv, int ql, const LfTView &tv, int tl, uint2 *planes, int32_t *col, int8_t *hb = nullptr)
{
    /* MULTI: queries of more than LANES * WPL words run as strips of that many words one after the other, chained through
     * hb[] (hout of a strip's bottom row per column) -- the junk heads / tails of 10+ kbp that wrong-candidate chains of
     * 20 kbp reads leave (configs[3]); planes are not stored in this form. */
    static_assert(!(MULTI && (FLAGS & LF_PASS_STORE)), "");
    constexpr int CB = LF_WAVE_CB;
    constexpr int SW = LANES * WPL;
    const int gl = (int)(threadIdx.x & (LANES - 1));
    const int n = (ql + 31) >> 5;
    const int S = MULTI ? (n + SW - 1) / SW : 1;
    const int wl = ql > 0 ? (ql - 1) >> 5 : -1;
    const uint32_t bl = (uint32_t)(ql - 1) & 31u;
    const int nblocks = (tl + CB - 1) / CB;
    int score = ql, best = ql, bestc = -1;     /* followed by the lane that owns row ql-1 */
    int colbase = tl;                          /* D(first row of the strip, tl) */
    if ((FLAGS & LF_PASS_COL) && gl == 0 && ql > 0) col[0] = tl;
    for (int s = 0; s < S; s++) {
    const int nvw = MULTI ? (n - s * SW < SW ? n - s * SW : SW) : n;
    const int nv = (nvw + WPL - 1) / WPL;
    const bool valid = gl < nv;
    const int w0 = (s * LANES + gl) * WPL;
    const int nsteps = lf_warp_max(ql > 0 ? nblocks + nv - 1 : 0);
    uint32_t lo[WPL], hi[WPL], nn[WPL], Pv[WPL], Mv[WPL], wm[WPL];
#pragma unroll
    for (int k = 0; k < WPL; k++) {
        lo[k] = 0; hi[k] = 0; nn[k] = 0xffffffffu;
        if (valid && w0 + k < n) lf_q32(d, qv, (int64_t)(w0 + k) * 32, lo[k], hi[k], nn[k]);
        Pv[k] = 0xffffffffu; Mv[k] = 0u;
        wm[k] = ((FLAGS & LF_PASS_SHW) && w0 + k == wl) ? 1u << bl : 0u;   /* the bit of row ql-1 */
    }
    LfTStream ts;
    ts.init(d.pac, ql > 0 ? tv.t0 : 0, ql > 0 ? tv.dir : 1);
    uint32_t pay = 0;                          /* (Ph, Mh) of this lane's bottom row for the columns of its last block */
    for (int step = 0; step < nsteps; step++) {
        uint32_t in = __shfl_up_sync(LF_FULL, pay, 1, LANES);
        const int cb = step - gl;
        pay = 0;
        if (valid && cb >= 0 && cb < nblocks) {
            const int cbase = cb * CB;
            const int ncol = tl - cbase < CB ? tl - cbase : CB;
            if (gl == 0) {
                in = 0x5555u;                  /* row 0 of a global alignment grows by one per column */
                if (MULTI && s > 0) {          /* ... or continues the strip above */
                    in = 0;
                    for (int ci = 0; ci < ncol; ci++) { const int h = hb[cbase + ci]; in |= ((h > 0 ? 1u : 0u) | (h < 0 ? 2u : 0u)) << (2 * ci); }
                }
            }
            uint32_t tb = ts.peek();
            ts.advance(ncol);
            uint2 *dst = (FLAGS & LF_PASS_STORE) ? planes + ((size_t)cbase * nv + gl) * WPL : nullptr;
            const bool feed = MULTI && gl == nv - 1 && s + 1 < S;   /* this lane's bottom row is the next strip's top */
            int sh2 = 0;
            for (int ci = 0; ci < ncol; ci++) {
                const uint32_t shi = (uint32_t)((int32_t)tb >> 31), slo = (uint32_t)((int32_t)(tb << 1) >> 31);
                tb <<= 2;
                const uint32_t hp = in & 1u, hm = (in >> 1) & 1u;   /* hin = +1 / -1 */
                in >>= 2;
                uint32_t nEq[WPL], a[WPL], sum[WPL];
#pragma unroll
                for (int k = 0; k < WPL; k++) {
                    nEq[k] = lf_neq(lo[k], hi[k], nn[k], slo, shi);
                    a[k] = Pv[k] & ~nEq[k];
                }
                lf_add_chain_cin<WPL>(a, Pv, sum, hm);             /* hin = -1 enters as the carry-in */
                uint32_t pPh = hp << 31, pMh = hm << 31;
                uint32_t phs = 0, mhs = 0;
#pragma unroll
                for (int k = 0; k < WPL; k++) {
                    const uint32_t Xh = (sum[k] ^ Pv[k]) | ~nEq[k] | (k == 0 ? hm : 0u);
                    const uint32_t Ph = Mv[k] | ~(Xh | Pv[k]);
                    const uint32_t Mh = Pv[k] & Xh;
                    const uint32_t Xv = ~nEq[k] | Mv[k];
                    if (FLAGS & LF_PASS_SHW) { phs |= Ph & wm[k]; mhs |= Mh & wm[k]; }
                    const uint32_t Phs = __funnelshift_l(pPh, Ph, 1), Mhs = __funnelshift_l(pMh, Mh, 1);
                    pPh = Ph; pMh = Mh;
                    const uint32_t nPv = Mhs | ~(Xv | Phs);
                    const uint32_t nMv = Phs & Xv;
                    if (FLAGS & LF_PASS_STORE) {
                        const uint32_t diagx = ~(nPv | Ph) & nEq[k];
                        dst[k] = make_uint2(nPv | diagx, (~nPv & Ph) | diagx);
                    }
                    Pv[k] = nPv; Mv[k] = nMv;
                }
                if (FLAGS & LF_PASS_SHW) {
                    score += (int)(phs != 0u) - (int)(mhs != 0u);
                    if (score < best) { best = score; bestc = cbase + ci; }
                }
                pay |= ((pPh >> 31) | ((pMh >> 31) << 1)) << sh2;
                sh2 += 2;
                if (feed) hb[cbase + ci] = (int8_t)((int)(pPh >> 31) - (int)(pMh >> 31));
                if (FLAGS & LF_PASS_STORE) dst += (size_t)nv * WPL;
            }
        }
    }
    /* last column: vertical deltas -> D(ql, tl), and D(x, tl) for every x if asked */
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < WPL; k++) {
        const int w = w0 + k;
        const uint32_t m = !valid ? 0u : w < wl ? 0xffffffffu : w == wl ? (0xffffffffu >> (31u - bl)) : 0u;
        cnt += __popc(Pv[k] & m) - __popc(Mv[k] & m);
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) { const int v = __shfl_up_sync(LF_FULL, incl, o, LANES); if (gl >= o) incl += v; }
    if (FLAGS & LF_PASS_COL) {
        if (valid) {
            int v = colbase + incl - cnt;
#pragma unroll
            for (int k = 0; k < WPL; k++) {
                const int w = w0 + k;
                const int rows = ql - w * 32 < 32 ? ql - w * 32 : 32;
                for (int b = 0; b < rows; b++) { v += (int)((Pv[k] >> b) & 1u) - (int)((Mv[k] >> b) & 1u); col[w * 32 + b + 1] = v; }
            }
        }
    }
    colbase += __shfl_sync(LF_FULL, incl, LANES - 1, LANES);
    if (MULTI) __syncwarp();                   /* hb[] written by the last lane is read by lane 0 of the next strip */
    }
    LfPassOut o;
    o.ed = colbase;
    const int owner = wl >= 0 ? (wl % SW) / WPL : 0;
    o.best = __shfl_sync(LF_FULL, best, owner, LANES);
    o.bestc = __shfl_sync(LF_FULL, bestc, owner, LANES);
    return o;
}

template <int WPL>
__device__ __forceinline__ LfPassOut lf_wave_pass_w(const LfDev &d, const LfQView &qv, int ql, const LfTView &tv, int tl, int flags,
                                                    uint2 *planes, int8_t *hb, int32_t *col)
{
    if constexpr (WPL == 8) if (((ql + 31) >> 5) > 32 * WPL) {   /* several strips (queries above 8192 rows) */
        switch (flags) {
        case LF_PASS_SHW: return lf_gwave<32, WPL, LF_PASS_SHW, true>(d, qv, ql, tv, tl, planes, col, hb);
        case LF_PASS_COL: return lf_gwave<32, WPL, LF_PASS_COL, true>(d, qv, ql, tv, tl, planes, col, hb);
        case 0: return lf_gwave<32, WPL, 0, true>(d, qv, ql, tv, tl, planes, col, hb);
        default: return lf_wave_pass_t<WPL>(d, qv, ql, tv, tl, flags, planes, hb, col);   /* planes of a leaf that tall: the general form */
        }
    }
    switch (flags) {
    case LF_PASS_STORE: return lf_gwave<32, WPL, LF_PASS_STORE>(d, qv, ql, tv, tl, planes, col);
    case LF_PASS_SHW: return lf_gwave<32, WPL, LF_PASS_SHW>(d, qv, ql, tv, tl, planes, col);
    case LF_PASS_COL: return lf_gwave<32, WPL, LF_PASS_COL>(d, qv, ql, tv, tl, planes, col);
    default: return lf_gwave<32, WPL, 0>(d, qv, ql, tv, tl, planes, col);
    }
}

#ifdef LF_EMU
#define LF_NOINLINE
#else
#define LF_NOINLINE __noinline__
#endif
/* one copy per kernel (k_myers_large calls it from eight places and it holds 20 instantiations of the pass) */
__device__ LF_NOINLINE LfPassOut lf_wave_pass(const LfDev &d, const LfQView &qv, int ql, const LfTView &tv, int tl, int flags,
                                              uint2 *planes, int8_t *hb, int32_t *col)
{
    switch (lf_wpl(ql)) {
    case 1: return lf_wave_pass_w<1>(d, qv, ql, tv, tl, flags, planes, hb, col);
    case 2: return lf_wave_pass_w<2>(d, qv, ql, tv, tl, flags, planes, hb, col);
    case 4: return lf_wave_pass_w<4>(d, qv, ql, tv, tl, flags, planes, hb, col);
    default: return lf_wave_pass_w<8>(d, qv, ql, tv, tl, flags, planes, hb, col);
    }
}

/* Canonical traceback over stored planes.  The walk is serial, so its cost is latency: the 32
 * lanes fetch the plane words of the next 32 columns of the current word-row at once and the walk
 * then reads them by shuffle instead of taking an L2 round trip per step.  Every lane walks the same
 * path; lane 0 writes one byte per op right-aligned below `hi_pos`.  Returns the number of ops. */
__device__ __forceinline__ int lf_large_traceback(const uint2 *planes, int ql, int tl, uint8_t *opsb, long long hi_pos)
{
    const int lane = threadIdx.x & 31;
    const int n = (ql + 31) >> 5;
    const int WPL = lf_wpl(ql), SW = 32 * WPL;
    long long p = hi_pos;
    int i = ql, j = tl;
    while (i > 0 && j > 0) {
        const int w = (i - 1) >> 5, s = w / SW, l = (w % SW) / WPL, k = w % WPL;
        const int nvw = n - s * SW < SW ? n - s * SW : SW;
        const int nv = (nvw + WPL - 1) / WPL;
        const unsigned long long sb = (unsigned long long)s * (unsigned long long)tl * 32ull * (unsigned long long)WPL;
        const int jt = j, colr = j - 1 - lane;
        uint2 v = make_uint2(0u, 0u);
        if (colr >= 0) v = planes[sb + ((unsigned long long)colr * nv + l) * WPL + k];
        while (i > 0 && j > 0 && ((i - 1) >> 5) == w && jt - j < 32) {
            const int kk = jt - j;
            const uint32_t x = __shfl_sync(LF_FULL, v.x, kk), y = __shfl_sync(LF_FULL, v.y, kk);
            const uint32_t b = (uint32_t)(i - 1) & 31u;
            const uint32_t op = ((x >> b) & 1u) | (((y >> b) & 1u) << 1);
            --p;
            if (lane == 0) opsb[p] = (uint8_t)op;
            i -= (op != 2u);
            j -= (op != 1u);
        }
    }
    while (i > 0) { --p; if (lane == 0) opsb[p] = 1; i--; }
    while (j > 0) { --p; if (lane == 0) opsb[p] = 2; j--; }
    __syncwarp();
    return (int)(hi_pos - p);
}

__device__ __forceinline__ void lf_warp_fill(uint8_t *dst, int n, uint8_t v)
{
    for (int k = threadIdx.x & 31; k < n; k += 32) dst[k] = v;
    __syncwarp();
}
__device__ __forceinline__ void lf_warp_move_down(uint8_t *buf, long long dst, long long src, int n)
{ /* dst <= src; chunks of 32 are read by every lane before any lane writes */
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < n; k += 32) {
        uint8_t v = (k + lane < n) ? buf[src + k + lane] : 0;
        __syncwarp();
        if (k + lane < n) buf[dst + k + lane] = v;
        __syncwarp();
    }
}

/* Scratch regions of one warp slot of k_myers_large */
struct LfLargeScr {
    uint2 *planes; int8_t *hb; int32_t *Lc, *Rc; uint8_t *opsb; int32_t *stack;
    __device__ __forceinline__ LfLargeScr(uint8_t *scr, const LfLargeCfg &cfg)
        : planes((uint2 *)scr), hb((int8_t *)(scr + cfg.off_hb)), Lc((int32_t *)(scr + cfg.off_L)), Rc((int32_t *)(scr + cfg.off_R)), opsb(scr + cfg.off_opsb),
          stack((int32_t *)(scr + cfg.off_stack)) { }
};

/* Split of (query [qo, qo+ql), target [to, to+tl)) at target column tl/2 (edlib.cpp:1176-1289): the smallest interior row x
 * with L[x] + R[x] == best, then the top boundary, then the bottom one.  Lc / Rc hold D(x, lw) of the left half and of the
 * reversed right half.  best < 0: not known yet -- the minimum over all rows is the distance of the piece (returned in
 * best).  Returns x, or -1 if no row fits (cannot happen for a correct distance).  One warp. */
__device__ __forceinline__ int lf_large_split_row(const int32_t *Lc, const int32_t *Rc, int ql, int lw, int rw, int &best, int &ls, int &rs)
{
    const int lane = threadIdx.x & 31;
    if (best < 0) {
        int m = 0x7fffffff;
        for (int x0 = lane; x0 <= ql; x0 += 32) { const int v = Lc[x0] + Rc[ql - x0]; m = v < m ? v : m; }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const int v = __shfl_xor_sync(LF_FULL, m, o); m = v < m ? v : m; }
        best = m;
    }
    int x = -1;
    for (int x0 = 1; x0 <= ql - 1 && x < 0; x0 += 32) {
        const int xx = x0 + lane;
        const bool hit = xx <= ql - 1 && Lc[xx] + Rc[ql - xx] == best;
        const uint32_t bal = __ballot_sync(LF_FULL, hit);
        if (bal) x = x0 + __ffs((int)bal) - 1;
    }
    if (x >= 0) { ls = Lc[x]; rs = Rc[ql - x]; }
    else if (lw + Rc[ql] == best) { x = 0; ls = lw; rs = Rc[ql]; }
    else if (Lc[ql] + rw == best) { x = ql; ls = Lc[ql]; rs = rw; }
    return x;
}

/* obtainAlignment (edlib.cpp:1090-1143) of (query [q0, q0+qlen), target [t0, t0+tlen)) of the task's views with an explicit
 * stack, by ONE warp in its own scratch: byte ops land in S.opsb[0, return value).  stored: the planes of the whole piece are
 * already in S.planes (a leaf whose store pass has run). */
__device__ __forceinline__ long long lf_large_path(const LfDev &d, const LfQView &qv, const LfTView &tv, int q0, int qlen, int t0, int tlen, int best0,
                                                   const LfLargeScr &S, bool stored, int &ed_io, int &status)
{
    const int lane = threadIdx.x & 31;
    long long outpos = 0;
    int sp = 0;
    if (lane == 0) { S.stack[0] = q0; S.stack[1] = qlen; S.stack[2] = t0; S.stack[3] = tlen; S.stack[4] = best0; }
    sp = 1;
    __syncwarp();
    bool first = true;
    while (sp > 0) {
        sp--;
        const int qo = S.stack[sp * 5 + 0], ql = S.stack[sp * 5 + 1], to = S.stack[sp * 5 + 2], tl = S.stack[sp * 5 + 3], best_in = S.stack[sp * 5 + 4];
        __syncwarp();
        if (ql == 0) { lf_warp_fill(S.opsb + outpos, tl, 2); outpos += tl; first = false; continue; }
        if (tl == 0) { lf_warp_fill(S.opsb + outpos, ql, 1); outpos += ql; first = false; continue; }
        if (lf_is_leaf((uint32_t)ql, (uint32_t)tl)) {
            if (!(stored && first)) {
                LfQView sq = lf_qsub(qv, qo, ql, false);
                LfTView st = lf_tsub(tv, to, tl, false);
                lf_wave_pass(d, sq, ql, st, tl, LF_PASS_STORE, S.planes, S.hb, nullptr);
                __syncwarp();
            }
            const long long hi_pos = (long long)(qo - q0) + (to - t0) + ql + tl;
            const int nops = lf_large_traceback(S.planes, ql, tl, S.opsb, hi_pos);
            lf_warp_move_down(S.opsb, outpos, hi_pos - nops, nops);
            outpos += nops;
            first = false;
            continue;
        }
        first = false;
        /* split at column tl/2 (edlib.cpp:1176-1196) */
        const int lw = tl / 2, rw = tl - lw;
        lf_wave_pass(d, lf_qsub(qv, qo, ql, false), ql, lf_tsub(tv, to, lw, false), lw, LF_PASS_COL, S.planes, S.hb, S.Lc);
        lf_wave_pass(d, lf_qsub(qv, qo, ql, true), ql, lf_tsub(tv, to + lw, rw, true), rw, LF_PASS_COL, S.planes, S.hb, S.Rc);
        __syncwarp();
        int best = best_in, ls = 0, rs = 0;
        const int x = lf_large_split_row(S.Lc, S.Rc, ql, lw, rw, best, ls, rs);
        if (best_in < 0) ed_io = best;   /* distance of the whole piece = min over all split rows */
        if (x < 0) { status = LF_ERR_CUDA; break; } /* cannot happen for a correct distance */
        __syncwarp();
        if (sp + 2 > LF_LARGE_STACK) { status = LF_ERR_NOMEM; break; }
        if (lane == 0) {
            S.stack[sp * 5 + 0] = qo + x; S.stack[sp * 5 + 1] = ql - x; S.stack[sp * 5 + 2] = to + lw; S.stack[sp * 5 + 3] = rw; S.stack[sp * 5 + 4] = rs;
            S.stack[sp * 5 + 5] = qo; S.stack[sp * 5 + 6] = x; S.stack[sp * 5 + 7] = to; S.stack[sp * 5 + 8] = lw; S.stack[sp * 5 + 9] = ls;
        }
        sp += 2;
        __syncwarp();
    }
    return outpos;
}

/* One task on a block of one or two warps (blockDim.x = 32 | 64).  With two warps a task above edlib's size rule is split
 * once at the top by both of them -- the left half pass on warp 0, the reversed right half pass on warp 1 -- and each warp
 * then aligns its own side of the split (lf_large_path); the op strings are joined when both are done.  The top-level
 * passes are half of a Hirschberg alignment's work and the two sides the other half, so a multi-kbp task (the
 * wrong-candidate gaps of configs[3]: one warp needed ~10 ms for 5 kbp x 5 kbp) takes about half as long.  Everything
 * else -- leaves, distance-only tasks -- runs on warp 0 alone. */
__device__ __forceinline__ void lf_large_task(const LfDev &d, uint32_t ti, const LfLargeCfg &cfg, uint8_t *scr0, uint8_t *scr1, int32_t *sh)
{
    const int lane = threadIdx.x & 31, warp = (int)(threadIdx.x >> 5);
    const bool two = blockDim.x > 32;
    const lf_align_task task = d.tasks[ti];
    const int q = (int)task.q_len, t = (int)task.t_len;
    LfQView qv; LfTView tv;
    lf_task_views(d, task, qv, tv);
    const LfLargeScr S0(scr0, cfg), S1(two ? scr1 : scr0, cfg);
    const LfLargeScr &S = warp == 0 ? S0 : S1;

    const bool shw = task.mode == LF_MODE_SHW;
    const bool want = !(task.flags & LF_F_NO_PATH);
    int ed = -1, end = t - 1;
    bool stored = false; /* planes of the whole task are already in `planes` */
    if (warp == 0) {
        if (!shw && want && lf_is_leaf((uint32_t)q, (uint32_t)t)) {
            LfPassOut o = lf_wave_pass(d, qv, q, tv, t, LF_PASS_STORE, S.planes, S.hb, nullptr);
            ed = o.ed; stored = true;
        } else if (!shw && want) {
            ed = -1; /* the first split yields min_x L[x]+R[x] = the distance */
        } else {
            LfPassOut o = lf_wave_pass(d, qv, q, tv, t, shw ? LF_PASS_SHW : 0, S.planes, S.hb, nullptr);
            if (shw) { ed = o.best; end = o.bestc; } else ed = o.ed;
        }
    }
    if (two) {   /* warp 1 needs the prefix-mode end column and the distance */
        if (warp == 0 && lane == 0) { sh[0] = ed; sh[1] = end; }
        __syncthreads();
        ed = sh[0]; end = sh[1];
        __syncthreads();
    }
    const uint64_t slot_hi = d.slot_end[ti] * 16ull;
    lf_align_result r;
    r.edit_distance = ed; r.end_location = end; r.status = 0; r.ops_len = 0; r.ops_off = slot_hi;
    if (!want) { if (warp == 0 && lane == 0) d.res[ti] = r; return; }

    const int teff = end + 1;
    int status = 0;
    long long n0 = 0, n1 = 0;
    const bool par = two && q > 0 && teff > 1 && !lf_is_leaf((uint32_t)q, (uint32_t)teff);
    if (!par) {
        if (warp == 0) n0 = lf_large_path(d, qv, tv, 0, q, 0, teff, ed, S0, stored, ed, status);
    } else {
        const int lw = teff / 2, rw = teff - lw;
        if (warp == 0) lf_wave_pass(d, qv, q, lf_tsub(tv, 0, lw, false), lw, LF_PASS_COL, S0.planes, S0.hb, S0.Lc);
        else lf_wave_pass(d, lf_qsub(qv, 0, q, true), q, lf_tsub(tv, lw, rw, true), rw, LF_PASS_COL, S1.planes, S1.hb, S0.Rc);
        __syncthreads();
        int best = ed, ls = 0, rs = 0;
        const int x = lf_large_split_row(S0.Lc, S0.Rc, q, lw, rw, best, ls, rs);   /* both warps, same answer */
        ed = best;
        __syncthreads();   /* S0.Lc / S0.Rc are free again */
        if (x < 0) status = LF_ERR_CUDA;
        else if (warp == 0) n0 = lf_large_path(d, qv, tv, 0, x, 0, lw, ls, S0, false, ls, status);
        else n1 = lf_large_path(d, qv, tv, x, q - x, lw, rw, rs, S1, false, rs, status);
    }
    if (two) {
        if (lane == 0) { sh[2 + warp] = (int32_t)(warp == 0 ? n0 : n1); sh[4 + warp] = status; }
        __syncthreads();
        n0 = sh[2]; n1 = sh[3]; status = sh[4] ? sh[4] : sh[5];
    }
    /* pack one byte per op into the 2-bit stream, left-aligned in the task's slot */
    const long long outpos = n0 + n1;
    const uint64_t slot_lo_w = d.slot_end[ti] - ((uint32_t)(q + t + 15) >> 4);
    const long long nwords = (outpos + 15) >> 4;
    for (long long wi = threadIdx.x; wi < nwords; wi += blockDim.x) {
        uint32_t word = 0;
        for (int k = 0; k < 16; k++) {
            const long long pp = wi * 16 + k;
            if (pp < outpos) word |= (uint32_t)(pp < n0 ? S0.opsb[pp] : S1.opsb[pp - n0]) << (2 * k);
        }
        d.ops[slot_lo_w + (uint64_t)wi] = word;
    }
    r.edit_distance = ed;
    r.ops_off = slot_lo_w * 16ull; r.ops_len = (uint32_t)outpos; r.status = status;
    if (warp == 0 && lane == 0) d.res[ti] = r;
    if (two) __syncthreads(); else __syncwarp();
}

__global__ void __launch_bounds__(64) k_myers_large(LfDev d, const uint32_t *__restrict__ order, uint32_t first, uint32_t count, LfLargeCfg cfg,
                                                    const uint32_t *__restrict__ count_ptr)
{   /* count_ptr != nullptr: `order + first` is a dense list whose length a previous kernel on the stream left there (the
     * tasks k_myers_bandreg could not certify: a warp finishes one of them in a fraction of the time a single thread
     * of the full-width kernel would, and that latency is the tail of the step) */
    __shared__ int32_t s_sh[8];
    __shared__ uint32_t s_k;
    const bool two = blockDim.x > 32;
    if (count_ptr) count = *count_ptr;
    uint8_t *scr0 = cfg.base + (unsigned long long)blockIdx.x * (two ? 2ull : 1ull) * cfg.stride;
    uint8_t *scr1 = scr0 + cfg.stride;
    for (;;) {
        uint32_t k = 0;
        if (two) {
            if (threadIdx.x == 0) s_k = atomicAdd(cfg.queue, 1u);
            __syncthreads();
            k = s_k;
            __syncthreads();
        } else {
            if (threadIdx.x == 0) k = atomicAdd(cfg.queue, 1u);
            k = __shfl_sync(LF_FULL, k, 0);
        }
        if (k >= count) break;
        lf_large_task(d, order[first + k], cfg, scr0, scr1, s_sh);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* k_myers_group: LANES lanes per task, 32 / LANES tasks per warp                                */
/* ------------------------------------------------------------------------------------------ */
/* Tasks too long for one thread (their latency would be the tail of the step) and too short for a whole warp (most
 * lanes would idle): prefix-mode heads / tails and far-off-diagonal gaps of 257 .. 2048 rows with their path (PATH:
 * leaves of edlib's size rule only, op planes in the group's scratch, canonical walk), and distance-only tasks of up to
 * 8192 rows (!PATH: the junk heads / tails the chain operator only needs the clip test for, src/LordFAST.cpp:1840,
 * :2175, and the forward half of the inversion test, :2037).  A warp takes 32 / LANES consecutive tasks of the sorted
 * class list at a time (same length bucket), runs them in lockstep through lf_gwave and walks the paths together. */
struct LfGroupRun { uint8_t *base; unsigned long long stride; uint32_t *queue; };

/* Canonical traceback (Up > Left > Diagonal, edlib.cpp:950, :984, :1015) of every group of the warp over its stored
 * planes.  The 32 columns left of the current position, current word-row, are fetched into the group's shared-memory
 * tile by its lanes together; every lane of the group then walks the same path out of the tile, lane 0 packs the ops
 * (16 per word) right-aligned below slot_hi.  Returns the number of ops. */
template <int LANES>
__device__ __forceinline__ uint32_t lf_group_walk(const uint2 *planes, int nvw, int ql, int j0, uint2 *tile, uint32_t *ops, uint64_t slot_hi, bool active)
{
    const int gl = (int)(threadIdx.x & (LANES - 1));
    int i = active ? ql : 0, j = active ? j0 : 0;
    uint32_t *wptr = ops + (slot_hi >> 4) - 1;
    uint32_t cur = 0, nops = 0;
    int sh = 30;
#define LF_EMIT(op) do { cur |= (uint32_t)(op) << sh; nops++; if (sh == 0) { if (gl == 0) *wptr = cur; wptr--; cur = 0; sh = 30; } else sh -= 2; } while (0)
    while (__any_sync(LF_FULL, i > 0 && j > 0)) {
        const bool go = i > 0 && j > 0;
        const int w = go ? (i - 1) >> 5 : 0;
        const int clo = j > 32 ? j - 32 : 0;
        if (go) for (int x = gl; x < j - clo; x += LANES) tile[x] = planes[(size_t)(clo + x) * nvw + w];
        __syncwarp();
        if (go) {
            const int rowlo = w * 32;
            while (i > rowlo && j > clo) {
                const uint2 v = tile[j - 1 - clo];
                const uint32_t b = (uint32_t)(i - 1) & 31u;
                const uint32_t op = ((v.x >> b) & 1u) | (((v.y >> b) & 1u) << 1);
                LF_EMIT(op);
                i -= (op != 2u);
                j -= (op != 1u);
            }
        }
        __syncwarp();
    }
    while (i > 0) { LF_EMIT(1u); i--; }
    while (j > 0) { LF_EMIT(2u); j--; }
    if (gl == 0 && sh != 30) *wptr = cur;
#undef LF_EMIT
    return nops;
}

template <int LANES, int WPL, bool PATH>
__global__ void __launch_bounds__(128) k_myers_group(LfDev d, const uint32_t *__restrict__ order, uint32_t first, uint32_t count, LfGroupRun cfg)
{
    constexpr int G = 32 / LANES;
    __shared__ uint2 s_tile[PATH ? 4 * G * 32 : 1];
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5), g = lane / LANES, gl = lane & (LANES - 1);
    const unsigned long long slot = ((unsigned long long)blockIdx.x * 4ull + (unsigned)warp) * (unsigned)G + (unsigned)g;
    uint2 *planes = PATH ? (uint2 *)(cfg.base + slot * cfg.stride) : nullptr;
    uint2 *tile = s_tile + (PATH ? (warp * G + g) * 32 : 0);
    for (;;) {
        uint32_t k0 = 0;
        if (lane == 0) k0 = atomicAdd(cfg.queue, (uint32_t)G);
        k0 = __shfl_sync(LF_FULL, k0, 0);
        if (k0 >= count) break;
        const bool have = k0 + (uint32_t)g < count;
        uint32_t ti = 0;
        lf_align_task task;
        task.read_id = 0; task.q_off = 0; task.q_len = 0; task.t_off = 0; task.t_len = 0; task.flags = 0; task.mode = 0; task.reserved = 0;
        LfQView qv; LfTView tv;
        qv.bit0 = 0; qv.dir = 1; qv.comp = 0; tv.t0 = 0; tv.dir = 1;
        if (have) { ti = order[first + k0 + (uint32_t)g]; task = d.tasks[ti]; lf_task_views(d, task, qv, tv); }
        const int q = (int)task.q_len, t = (int)task.t_len;
        const bool shw = have && task.mode == LF_MODE_SHW;
        LfPassOut o;
        if (__any_sync(LF_FULL, shw)) o = lf_gwave<LANES, WPL, (PATH ? LF_PASS_STORE : 0) | LF_PASS_SHW>(d, qv, q, tv, t, planes, nullptr);
        else o = lf_gwave<LANES, WPL, (PATH ? LF_PASS_STORE : 0)>(d, qv, q, tv, t, planes, nullptr);
        const int ed = shw ? o.best : o.ed, end = shw ? o.bestc : t - 1;
        const uint64_t slot_hi = have ? d.slot_end[ti] * 16ull : 0ull;
        uint32_t nops = 0;
        if (PATH) {
            const bool want = have && !(task.flags & LF_F_NO_PATH);
            const int nvw = (((q + 31) >> 5) + WPL - 1) / WPL * WPL;
            __syncwarp();
            nops = lf_group_walk<LANES>(planes, nvw, q, end + 1, tile, d.ops, slot_hi, want);
        }
        if (have && gl == 0) {
            lf_align_result r;
            r.edit_distance = ed; r.end_location = end; r.status = 0; r.ops_off = slot_hi - nops; r.ops_len = nops;
            d.res[ti] = r;
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------ */
/* k_ksw_extend: ksw_extend2 (lib/bwa/ksw.c:380-479), one thread per task                      */
/* ------------------------------------------------------------------------------------------ */
struct LfExtDev {
    const uint8_t *pac; int64_t l_pac;
    const uint8_t *bases; const uint64_t *read_off; uint32_t n_reads;
    const lf_extend_task *tasks; uint32_t n_tasks;
    lf_extend_result *res;
    const uint64_t *scr_off; /* exclusive scan of (q_len+1) */
    int2 *scratch;           /* {H diagonal feed, E} per query column */
};

__device__ __forceinline__ uint32_t lf_char2int(uint32_t c)
{ /* src/LordFAST.cpp:158-164 */
    return (c == 'A' || c == 'a') ? 0u : (c == 'C' || c == 'c') ? 1u : (c == 'G' || c == 'g') ? 2u : (c == 'T' || c == 't') ? 3u : 4u;
}

__global__ void k_extend_prep(LfExtDev d, uint32_t *scr_items)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_tasks) return;
    lf_extend_task t = d.tasks[i];
    bool ok = t.q_len >= 1 && t.t_len >= 1 && t.read_id < d.n_reads && t.h0 > 0 && t.e_del > 0 && t.e_ins > 0;
    if (ok) {
        uint64_t L = d.read_off[t.read_id + 1] - d.read_off[t.read_id];
        ok = (uint64_t)t.q_off + t.q_len <= L && (int64_t)t.t_off + (int64_t)t.t_len <= d.l_pac;
    }
    scr_items[i] = ok ? t.q_len + 1u + (t.q_len + 7u) / 8u + 1u : 0u; /* eh[] + one query code byte per column */
}

/* One warp per task.  The reference's row loop is kept row by row (band re-trim, m==0 and z-drop exits
 * are row-sequential), but inside a row the 32 lanes take 32 consecutive columns at a time: E and the
 * diagonal feed come from the previous row, and F(i,j+1) = max(F(i,j)-e_ins, max(M-oe_ins,0)) does not
 * depend on H, so with u = F + j*e_ins it is a prefix maximum (5 shuffles per 32 columns).  Cells outside
 * [beg,end] keep their old contents exactly as in the reference (it reads such stale cells when the band
 * widens again).  eh[] is a ring of LF_KSW_RING columns in shared memory when the band fits (2w+2 <= ring: a
 * column and the column one ring length later are then never live together; columns are initialised with the
 * first-row values, ksw.c:395-397, when the band first reaches them), else the task's global scratch. */
#define LF_KSW_RING 256
/* Row loop of ksw_extend2 (ksw.c:416-470) with C consecutive columns per lane; eh[] is the shared-memory ring.  Same
 * arithmetic, band, exits and stale-cell behaviour as the generic loop in k_ksw_extend, cell for cell. */
template <int C, typename FirstRow>
__device__ __forceinline__ void lf_ksw_rows_blk(int2 *eh, const uint8_t *qcode, LfTCursor &tcur, int qlen, int tlen, int w, int h0, int o_del, int e_del,
                                                int oe_del, int oe_ins, int e_ins, int zdrop, int smatch, int smis, const FirstRow &first_row,
                                                int &best, int &best_i, int &best_j)
{
    const int lane = (int)(threadIdx.x & 31);
    const uint32_t jm = (uint32_t)(LF_KSW_RING - 1);
    const int NEG = -(1 << 29);
    int beg = 0, end = qlen, hi_init = 0;
    for (int r = 0; r < tlen; r++) {
        const uint32_t tc = tcur.next();
        if (beg < r - w) beg = r - w;
        if (end > r + w + 1) end = r + w + 1;
        if (end > qlen) end = qlen;
        if (hi_init < end) {
            for (int j = hi_init + lane; j < end; j += 32) eh[(uint32_t)j & jm] = first_row(j);
            hi_init = end;
            __syncwarp();
        }
        int h1 = 0;
        if (beg == 0) { h1 = h0 - (o_del + e_del * (r + 1)); if (h1 < 0) h1 = 0; }
        const int width = end - beg;            /* <= 2w + 1 <= 32 C */
        const int rr0 = lane * C;
        int M[C], e[C], g[C];
        int agg = NEG;                          /* max over the lane's cells of g + (rr + 1) e_ins */
#pragma unroll
        for (int k = 0; k < C; k++) {
            const int rr = rr0 + k, j = beg + rr;
            M[k] = 0; e[k] = 0; g[k] = 0;
            if (rr < width) {
                const int2 c = eh[(uint32_t)j & jm];
                const uint32_t qc = qcode[j];
                const int sc = qc > 3u ? 0 : (qc == tc ? smatch : smis);
                M[k] = c.x ? c.x + sc : 0;
                e[k] = c.y;
                int t = M[k] - oe_ins; t = t > 0 ? t : 0;
                g[k] = t;
                const int v = t + (rr + 1) * e_ins;
                agg = v > agg ? v : agg;
            }
        }
        int incl = agg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(LF_FULL, incl, o); if (lane >= o && x > incl) incl = x; }
        int excl = __shfl_up_sync(LF_FULL, incl, 1);
        int f = lane == 0 ? 0 : excl - rr0 * e_ins;   /* F of the lane's first cell (excl >= rr0 e_ins whenever a cell precedes it) */
        if (lane != 0 && excl == NEG) f = 0;
        int h[C], en[C];
        int lmax = -1, lmax_k = 0;
        uint32_t nzmask = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
            const int rr = rr0 + k;
            int hh = M[k] > e[k] ? M[k] : e[k];
            hh = hh > f ? hh : f;
            h[k] = hh;
            if (rr < width) {
                if (hh >= lmax) { lmax = hh; lmax_k = k; }
                int t = M[k] - oe_del; t = t > 0 ? t : 0;
                int x = e[k] - e_del; x = x > t ? x : t;
                en[k] = x;
            } else en[k] = 0;
            const int fn = f - e_ins;
            f = fn > g[k] ? fn : g[k];
        }
        /* H of the previous column of this row, for the diagonal feed of the next row */
        int nact = width - rr0; nact = nact < 0 ? 0 : nact > C ? C : nact;   /* the lane's active cells */
        int mylast = h1;
#pragma unroll
        for (int k = 0; k < C; k++) if (k == nact - 1) mylast = h[k];
        int prevlast = __shfl_up_sync(LF_FULL, mylast, 1);
        if (lane == 0) prevlast = h1;
#pragma unroll
        for (int k = 0; k < C; k++) {
            const int rr = rr0 + k, j = beg + rr;
            const int hprev = k == 0 ? prevlast : h[k - 1];
            if (rr < width) {
                eh[(uint32_t)j & jm] = int2{hprev, en[k]};
                if (hprev != 0 || en[k] != 0) nzmask |= 1u << k;
            }
        }
        /* row maximum, ties to the later column (ksw.c:437); cells that stay non-zero (ksw.c:466-469) */
        const int rowmax_all = lf_warp_max(lmax);
        int rowmax = 0, rowmax_j = -1;
        if (rowmax_all >= 0) {
            const uint32_t bal = __ballot_sync(LF_FULL, nact > 0 && lmax == rowmax_all);
            const int src = 31 - __clz((int)bal);
            rowmax = rowmax_all;
            rowmax_j = beg + __shfl_sync(LF_FULL, rr0 + lmax_k, src);
        }
        const uint32_t nzb = __ballot_sync(LF_FULL, nzmask != 0u);
        int first_nz = -1, last_nz = -1;
        if (nzb) {
            const int lf = __ffs((int)nzb) - 1, ll = 31 - __clz((int)nzb);
            const int kf = __ffs((int)nzmask) - 1, kl = 31 - __clz((int)nzmask);
            first_nz = beg + __shfl_sync(LF_FULL, rr0 + kf, lf);
            last_nz = beg + __shfl_sync(LF_FULL, rr0 + kl, ll);
        }
        const int h_last = __shfl_sync(LF_FULL, mylast, width > 0 ? (width - 1) / C : 0);
        if (lane == 0) eh[(uint32_t)end & jm] = int2{width > 0 ? h_last : h1, 0};
        if (hi_init < end + 1) hi_init = end + 1;
        __syncwarp();
        if (rowmax == 0) break;
        if (rowmax > best) { best = rowmax; best_i = r; best_j = rowmax_j; }
        else if (zdrop > 0) {
            const int di = r - best_i, dj = rowmax_j - best_j;
            if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
            else { if (best - rowmax - (dj - di) * e_ins > zdrop) break; }
        }
        const int hl = width > 0 ? h_last : h1;
        if (hl != 0) last_nz = end;
        const int nbeg = first_nz >= 0 ? first_nz : end;
        const int jl = last_nz >= nbeg ? last_nz : nbeg - 1;
        beg = nbeg;
        end = jl + 2 < qlen ? jl + 2 : qlen;
    }
}

__global__ void __launch_bounds__(128) k_ksw_extend(LfExtDev d)
{
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= d.n_tasks) return;
    const lf_extend_task t = d.tasks[i];
    lf_extend_result out; out.score = -1; out.qle = 0; out.tle = 0;
    const uint64_t so = d.scr_off[i];
    if (d.scr_off[i + 1] == so) { if (lane == 0) d.res[i] = out; return; } /* rejected by k_extend_prep */
    __shared__ int2 s_ring[4][LF_KSW_RING];
    int2 *eh = d.scratch + so;
    const int qlen = (int)t.q_len, tlen = (int)t.t_len;
    uint8_t *qcode = (uint8_t *)(eh + qlen + 1);
    const uint64_t ro = d.read_off[t.read_id];
    const int64_t L = (int64_t)(d.read_off[t.read_id + 1] - ro);
    const int rev1 = (t.flags & LF_F_REVERSE_BOTH) != 0;
    int64_t f0 = rev1 ? (int64_t)t.q_off + qlen - 1 : (int64_t)t.q_off;
    int qdir = rev1 ? -1 : 1;
    uint32_t comp = 0;
    if (t.flags & LF_F_READ_REV) { f0 = L - 1 - f0; qdir = -qdir; comp = 1; }
    const int64_t t0 = rev1 ? (int64_t)t.t_off + tlen - 1 : (int64_t)t.t_off;
    const int tdir = rev1 ? -1 : 1;
    const int smatch = 2, smis = t.matrix == LF_MAT_DEFAULT ? -5 : -16;
    const int o_del = t.o_del, e_del = t.e_del, o_ins = t.o_ins, e_ins = t.e_ins, h0 = t.h0, zdrop = t.zdrop;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int w = t.w;
    {   /* band clamp (ksw.c:399-407), max matrix entry is the match score */
        int lim = (int)((double)(qlen * smatch - o_ins) / e_ins + 1.);
        lim = lim > 1 ? lim : 1; w = w < lim ? w : lim;
        lim = (int)((double)(qlen * smatch - o_del) / e_del + 1.);
        lim = lim > 1 ? lim : 1; w = w < lim ? w : lim;
    }
    const bool ring = 2 * w + 2 <= LF_KSW_RING;
    const uint32_t jm = ring ? (uint32_t)(LF_KSW_RING - 1) : 0xffffffffu;   /* column -> cell index */
    if (ring) eh = s_ring[threadIdx.x >> 5];
    auto first_row = [&](int j) {   /* ksw.c:395-397 */
        int hv;
        if (j == 0) hv = h0;
        else if (j == 1) hv = h0 > oe_ins ? h0 - oe_ins : 0;
        else { const int prev = h0 - oe_ins - (j - 2) * e_ins; hv = (h0 > oe_ins && prev > e_ins) ? prev - e_ins : 0; }
        return int2{hv, 0};
    };
    /* query codes (src/LordFAST.cpp:158-164, 1191-1201) and, without the ring, the whole first row */
    for (int j = lane; j <= qlen; j += 32) {
        if (j < qlen) {
            uint32_t qc = lf_char2int(d.bases[ro + (uint64_t)(f0 + (int64_t)qdir * j)]);
            if (comp && qc < 4u) qc = 3u - qc;
            qcode[j] = (uint8_t)qc;
        }
        if (!ring) eh[j] = first_row(j);
    }
    int hi_init = 0;   /* ring: columns below hi_init hold first-row or later values */
    __syncwarp();
    const int NEG = -(1 << 29);
    int best = h0, best_i = -1, best_j = -1, beg = 0, end = qlen;
    LfTCursor tcur;
    tcur.init(d.pac, t0, tdir);
    /* Bands of up to 96 / 224 columns (the clip and split parameter sets: w = 40 / 100): every lane owns C consecutive
     * columns of the row, so a row is one round -- one prefix-max scan, one row-maximum reduction, two ballots -- instead
     * of one round per 32 columns; the per-task latency of a junk tail (as many rows as the tail is long) is what the
     * chain operator's round 2 waits for. */
    const int bandw = 2 * w + 1;
    if (ring && bandw <= 96) lf_ksw_rows_blk<3>(eh, qcode, tcur, qlen, tlen, w, h0, o_del, e_del, oe_del, oe_ins, e_ins, zdrop, smatch, smis, first_row, best, best_i, best_j);
    else if (ring && bandw <= 224) lf_ksw_rows_blk<7>(eh, qcode, tcur, qlen, tlen, w, h0, o_del, e_del, oe_del, oe_ins, e_ins, zdrop, smatch, smis, first_row, best, best_i, best_j);
    else {
    for (int r = 0; r < tlen; r++) {
            const uint32_t tc = tcur.next();
            if (beg < r - w) beg = r - w;
            if (end > r + w + 1) end = r + w + 1;
            if (end > qlen) end = qlen;
            if (ring && hi_init < end) {
                for (int j = hi_init + lane; j < end; j += 32) eh[(uint32_t)j & jm] = first_row(j);
                hi_init = end;
                __syncwarp();
            }
            int h1 = 0;
            if (beg == 0) { h1 = h0 - (o_del + e_del * (r + 1)); if (h1 < 0) h1 = 0; }
            int carry_u = 0;            /* u = F + (j-beg)*e_ins at the start of the round; F(i,beg) = 0 */
            int prev_h = h1;            /* H(i, j-1) for the first lane of the round */
            int rowmax = 0, rowmax_j = -1, first_nz = -1, last_nz = -1, h_last = h1;
            const int width = end - beg;
            for (int k0 = 0; k0 < width; k0 += 32) {
                const int rr = k0 + lane, j = beg + rr;
                const bool act = rr < width;
                int M = 0, e = 0;
                if (act) {
                    const int2 c = eh[(uint32_t)j & jm];
                    const uint32_t qc = qcode[j];
                    const int sc = qc > 3u ? 0 : (qc == tc ? smatch : smis);
                    M = c.x ? c.x + sc : 0;
                    e = c.y;
                }
                int g = M - oe_ins; g = g > 0 ? g : 0;
                const int v = act ? g + (rr + 1) * e_ins : NEG;
                int incl = v;
    #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(LF_FULL, incl, o); if (lane >= o && x > incl) incl = x; }
                int excl = __shfl_up_sync(LF_FULL, incl, 1);
                if (lane == 0) excl = NEG;
                const int u = carry_u > excl ? carry_u : excl;
                const int f = u - rr * e_ins;
                const int tot = __shfl_sync(LF_FULL, incl, 31);
                carry_u = carry_u > tot ? carry_u : tot;
                int h = M > e ? M : e;
                h = h > f ? h : f;
                if (!act) h = -1;
                int hprev = __shfl_up_sync(LF_FULL, h, 1);
                if (lane == 0) hprev = prev_h;
                int en = 0;
                if (act) {
                    int tt = M - oe_del; tt = tt > 0 ? tt : 0;
                    en = e - e_del; en = en > tt ? en : tt;
                    eh[(uint32_t)j & jm] = int2{hprev, en};
                }
                /* row maximum, ties to the later column (ksw.c:437) */
                const int mk = lf_warp_max(h);
                if (mk >= rowmax && mk >= 0) {
                    const uint32_t bal = __ballot_sync(LF_FULL, act && h == mk);
                    rowmax_j = beg + k0 + (31 - __clz((int)bal));
                    rowmax = mk;
                }
                /* cells that stay non-zero, for the band re-trim (ksw.c:466-469) */
                const uint32_t nzb = __ballot_sync(LF_FULL, act && (hprev != 0 || en != 0));
                if (nzb) {
                    if (first_nz < 0) first_nz = beg + k0 + __ffs((int)nzb) - 1;
                    last_nz = beg + k0 + (31 - __clz((int)nzb));
                }
                const int nact = width - k0 < 32 ? width - k0 : 32;
                h_last = __shfl_sync(LF_FULL, h, nact - 1);
                prev_h = h_last;
            }
            if (lane == 0) eh[(uint32_t)end & jm] = int2{h_last, 0};
            if (hi_init < end + 1) hi_init = end + 1;
            __syncwarp();
            if (rowmax == 0) break;
            if (rowmax > best) { best = rowmax; best_i = r; best_j = rowmax_j; }
            else if (zdrop > 0) {
                const int di = r - best_i, dj = rowmax_j - best_j;
                if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
                else { if (best - rowmax - (dj - di) * e_ins > zdrop) break; }
            }
            if (h_last != 0) last_nz = end;
            const int nbeg = first_nz >= 0 ? first_nz : end;
            const int jl = last_nz >= nbeg ? last_nz : nbeg - 1;
            beg = nbeg;
            end = jl + 2 < qlen ? jl + 2 : qlen;
        }
    }
    out.score = best; out.qle = best_j + 1; out.tle = best_i + 1;
    if (lane == 0) d.res[i] = out;
}

/* ------------------------------------------------------------------------------------------ */
/* INT32 issue-rate microbenchmark (roofline denominator; MEASURED_PEAKS.json has no INT32 figure) */
/* ------------------------------------------------------------------------------------------ */
template <int WHICH>
__global__ void __launch_bounds__(256) k_int32_peak(uint32_t *out, int iters, uint32_t seed)
{ /* 8 registers updated round-robin, every instruction with three distinct live register inputs so
   * that ptxas can neither fold two of them into one LOP3/IADD3 nor strength-reduce the loop; the
   * committed SASS (profiles/) shows 64 LOP3 / IADD3 / IMAD per iteration. */
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed * (uint32_t)(k + 3) + threadIdx.x * 8u + (uint32_t)k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t x = a[(k + 3) & 7], y = a[(k + 5) & 7];
                const bool lop = WHICH == 0 || ((WHICH == 2 || WHICH == 3) && (k & 1));
                if (lop) {
#if defined(__CUDA_ARCH__)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(a[k]) : "r"(x), "r"(y)); /* a ^ (x & y): one LOP3 */
#else
                    a[k] = a[k] ^ (x & y);
#endif
                } else if (WHICH == 1 || WHICH == 2) a[k] = a[k] + x + y;                        /* IADD3 */
                else a[k] = a[k] * x + y;                                                      /* IMAD  */
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) x ^= a[k];
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x; /* keeps the loop alive */
}

/* ------------------------------------------------------------------------------------------ */
/* k_emit_slots: the reference's CIGAR / MD accumulation on the GPU        */
/* ------------------------------------------------------------------------------------------ */
/* alignChain_edlib (src/LordFAST.cpp:1820-2249) appends, per chain: the head alignment, then for every
 * anchor its match run and the alignment of the gap behind it, then the tail, into run-length CIGAR
 * (edlibCigar_toString :1596-1626) and MD (edlibMD_toString :1717-1763) text.  That is a serial walk of
 * ~12 k ops per 10 kbp chain; here it is cut into SLOTS -- slot 0 = head, slot k = anchor k-1 and the gap
 * behind it, slot n = last anchor + tail -- with one block per chain and one thread per slot.
 *
 * What couples neighbouring slots is small, because every slot but the head starts with an anchor's match
 * run: (a) a run of 'M' left open at the end of a slot merges with that match run, so its length is handed on
 * (carry_c), and (b) the matches counted since the last MD event are handed on the same way (carry_m).  A slot
 * that consists of matches only passes both through.  Only the digits of the first number a slot writes depend
 * on the carry.  Two kernels, same walk:
 *   k_emit_slots<false>  walks every slot with carry 0, resolves the carries inside the block, sizes every
 *                        slot's text, scans the sizes, and stores (carry_c, carry_m, off_c, off_m, task) per slot
 *                        plus records / bytes per chain;
 *   (host: three scans over the per-chain totals place the chains)
 *   k_emit_slots<true>   walks again with the stored carries and writes the text in place; thread 0 then
 *                        assembles the chain's records (a split gap closes one record and opens the next;
 *                        records with fewer than two anchors are dropped, :1991 / :2063).
 * CIGARs and MDs go to two regions of one text buffer. */
struct LfSplitDev {
    uint32_t gap_i;                                  /* index of the seed before the gap, inside its chain */
    int32_t t_first, t_mid_r, t_second;              /* round-3 task indices or -1 */
    uint32_t qs2, ts2, qe2, te2;
    uint32_t split, inv;                             /* extensions did not cross / inversion accepted (host decides, double math) */
};
struct LfSlotInfo { uint32_t carry_c, carry_m, off_c, off_m; };
struct LfEmitDev {
    const uint8_t *pac;
    const lf_chain *chains; const lf_seed *seeds; uint32_t n_chains;
    const uint32_t *chain_list;    /* block b works on chain chain_list[b] (nullptr: chain b); per-list arrays are indexed by b */
    const uint64_t *read_off;
    const uint64_t *task_base;     /* round-1 task index of the chain's first task */
    const uint64_t *slot_base;     /* sum of (n_seeds + 1) over the chains before this one */
    const uint8_t *guards;         /* bit 0: head aligned, bit 1: tail aligned (else soft clip) */
    const int32_t *clip;           /* 4 per chain: head t3, head qle, tail t3, tail qle (t3 < 0: keep the round-1 alignment); nullptr: none */
    const uint32_t *split_begin;   /* n_chains + 1; nullptr: no chain of the list has a split */
    const LfSplitDev *splits;
    const lf_align_result *r1; const uint32_t *ops1;
    const lf_align_result *r3; const uint32_t *ops3;
    LfSlotInfo *slot_info; uint32_t *slot_task;   /* per slot, written by the sizing pass */
    /* sizing pass out, per list entry */
    uint32_t *nrec, *cig_bytes, *md_bytes;
    /* writing pass in/out; rec_off / cig_off / md_off are exclusive scans over the list with the total last */
    const uint64_t *rec_off, *cig_off, *md_off;
    lf_sam_record *recs; char *text;   /* this list's records and text: all CIGARs, then all MDs */
    uint64_t out_base;                 /* offset of text[0] in the caller-visible text buffer (goes into the records) */
};

__device__ __forceinline__ uint32_t lf_ndigits(uint32_t u)
{
    return u < 10u ? 1u : u < 100u ? 2u : u < 1000u ? 3u : u < 10000u ? 4u : u < 100000u ? 5u : u < 1000000u ? 6u : u < 10000000u ? 7u : u < 100000000u ? 8u : u < 1000000000u ? 9u : 10u;
}

/* Text builder of one slot.  WRITE = false only counts, and leaves out the digits of the first CIGAR number
 * and of the first MD number when they depend on the carry (deferred: cfirst / mfirst hold the local part). */
template <bool WRITE>
struct LfSlotB {
    char *cp, *mp;
    uint32_t cn, mn;
    char cch, mlast;
    uint32_t cnum, cnops, mnum;
    bool cdef, mdef;
    uint32_t cfirst, mfirst;
    __device__ __forceinline__ void start(bool fresh, uint32_t carry_c, uint32_t carry_m)
    {
        mlast = '='; cfirst = 0; mfirst = 0;
        if (fresh) { cch = 0; cnum = 0; cnops = 0; mnum = 0; cdef = false; mdef = false; }
        else { cch = 'M'; cnum = carry_c; cnops = 1; mnum = carry_m; cdef = !WRITE; mdef = !WRITE; }
    }
    __device__ __forceinline__ void putc_c(char c) { if (WRITE) *cp++ = c; cn++; }
    __device__ __forceinline__ void putc_m(char c) { if (WRITE) *mp++ = c; mn++; }
    __device__ __forceinline__ void num_c(uint32_t u)
    {
        if (!WRITE) { if (cdef) { cfirst = u; cdef = false; } else cn += lf_ndigits(u); return; }
        char t[12]; int n = 0;
        do { t[n++] = (char)('0' + u % 10u); u /= 10u; } while (u);
        while (n) putc_c(t[--n]);
    }
    __device__ __forceinline__ void num_m(uint32_t u)
    {
        if (!WRITE) { if (mdef) { mfirst = u; mdef = false; } else mn += lf_ndigits(u); return; }
        char t[12]; int n = 0;
        do { t[n++] = (char)('0' + u % 10u); u /= 10u; } while (u);
        while (n) putc_m(t[--n]);
    }
    __device__ __forceinline__ void flush_c(bool last)
    {
        if (!cnum) return;
        num_c(cnum);
        putc_c(((last || cnops == 0) && cch == 'I') ? 'S' : cch);
        cnops++; cnum = 0;
    }
    __device__ __forceinline__ void cig_run(char c, int n)
    {
        if (n <= 0) return;
        if (c == cch) { cnum += (uint32_t)n; return; }
        flush_c(false);
        cch = c; cnum = (uint32_t)n;
    }
    __device__ __forceinline__ void md_match(int n) { if (n > 0) { mnum += (uint32_t)n; mlast = '='; } }
    __device__ __forceinline__ void md_ins(int n) { if (n > 0) mlast = 'I'; }
    __device__ __forceinline__ void md_mismatch(char b) { num_m(mnum); mnum = 0; putc_m(b); mlast = 'X'; }
    __device__ __forceinline__ void md_del(char b) { if (mlast != 'D') { num_m(mnum); mnum = 0; putc_m('^'); } putc_m(b); mlast = 'D'; }
    __device__ __forceinline__ void run(char c, int n) { cig_run(c, n); if (c == 'I') md_ins(n); else md_match(n); }
    /* end of a slot inside a record: a trailing I / D run cannot merge with the anchor that follows */
    __device__ __forceinline__ void end_slot() { if (cch != 'M') { flush_c(false); cch = 0; } }
    /* end of a record */
    __device__ __forceinline__ void finish()
    {
        flush_c(true);
        num_m(mnum);
        putc_c('\0'); putc_m('\0');
    }
};

__device__ __forceinline__ uint32_t lf_op_at(const uint32_t *ops, uint64_t p) { return (ops[p >> 4] >> (((uint32_t)p & 15u) << 1)) & 3u; }
__device__ __forceinline__ char lf_pac_char(const uint8_t *pac, uint32_t l) { const uint32_t c = lf_tsym(pac, (int64_t)l); return c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : 'T'; }

/* ops of one alignment appended to the record; reversed = the task ran right-to-left; t0 = forward
 * reference position of the first target base covered.  Runs of matches are skipped a word at a time. */
template <class BT>
__device__ __forceinline__ void lf_emit_segment(BT &B, const uint32_t *ops, const lf_align_result &r, bool reversed, const uint8_t *pac, uint32_t t0)
{
    uint32_t tp = t0, k = 0;
    const uint32_t n = r.ops_len;
    while (k < n) {
        int run = 0;
        for (;;) {
            if (k >= n) break;
            uint32_t z, avail;
            if (!reversed) {
                const uint64_t p = r.ops_off + k;
                const uint32_t sh = ((uint32_t)p & 15u) << 1;
                const uint32_t w = ops[p >> 4] >> sh;
                avail = 16u - ((uint32_t)p & 15u);
                if (avail > n - k) avail = n - k;
                z = w ? (uint32_t)(__ffs((int)w) - 1) >> 1 : 16u;
            } else {
                const uint64_t p = r.ops_off + (n - 1 - k);
                const uint32_t top = (uint32_t)p & 15u;
                const uint32_t w = ops[p >> 4] << (30u - 2u * top);
                avail = top + 1u;
                if (avail > n - k) avail = n - k;
                z = w ? (uint32_t)__clz((int)w) >> 1 : 16u;
            }
            if (z >= avail) { run += (int)avail; k += avail; continue; }
            run += (int)z; k += z;
            break;
        }
        if (run) { B.cig_run('M', run); B.md_match(run); tp += (uint32_t)run; }
        if (k >= n) break;
        const uint64_t p = reversed ? r.ops_off + (n - 1 - k) : r.ops_off + k;
        const uint32_t op = lf_op_at(ops, p);
        k++;
        if (op == 1u) { B.cig_run('I', 1); B.md_ins(1); }
        else if (op == 2u) { B.cig_run('D', 1); B.md_del(lf_pac_char(pac, tp)); tp++; }
        else { B.cig_run('M', 1); B.md_mismatch(lf_pac_char(pac, tp)); tp++; }
    }
}

/* The accepted-inversion record: the reference appends the trailing clip to the END of the CIGAR deque but
 * to the BEGINNING of the MD deque (:2056-2057), so MD and CIGAR positions pair up out of step.  Emulated
 * index by index: cigar = I^a ops I^b, md = '-'^b '-'^a mdops. */
template <class BT>
__device__ __forceinline__ void lf_emit_inversion(BT &B, const uint32_t *ops, const lf_align_result &r, const uint8_t *pac, uint32_t t0, uint32_t a, uint32_t b)
{
    const uint32_t n = r.ops_len, total = a + n + b;
    /* CIGAR: plain run-length of the cigar chars */
    B.cig_run('I', (int)a);
    for (uint32_t k = 0; k < n; k++) { const uint32_t op = lf_op_at(ops, r.ops_off + k); B.cig_run(op == 1u ? 'I' : op == 2u ? 'D' : 'M', 1); }
    B.cig_run('I', (int)b);
    /* MD from the (md[i], cigar[i]) pairs */
    uint32_t tp = t0;
    for (uint32_t i = a + b; i < total; i++) {
        const uint32_t opm = lf_op_at(ops, r.ops_off + (i - a - b));      /* op whose MD char sits at position i */
        char m;
        if (opm == 0u) { m = '='; tp++; } else if (opm == 1u) m = '-'; else { m = lf_pac_char(pac, tp); tp++; }
        char c;                                                          /* cigar char at position i */
        if (i < a) c = 'I';
        else if (i < a + n) { const uint32_t opc = lf_op_at(ops, r.ops_off + (i - a)); c = opc == 1u ? 'I' : opc == 2u ? 'D' : 'M'; }
        else c = 'I';
        if (m == '=') { B.mnum++; B.mlast = '='; }
        else if (m == '-') { B.mlast = 'I'; }
        else if (c == 'M') { B.md_mismatch(m); }
        else if (c == 'D') { B.md_del(m); }
    }
    if (a + b > 0) { /* the leading '-' entries only set last = 'I' when nothing followed them */ if (total == a + b) B.mlast = 'I'; }
}

#define LF_EMIT_BLOCK 128
#define LF_EMIT_STAGE 12288u   /* bytes of CIGAR (and of MD) text one round of slots may stage in shared memory */
enum { LF_SL_HEAD = 1, LF_SL_PUSH = 2, LF_SL_INV = 4, LF_SL_SPLIT = 8, LF_SL_FINAL = 16 };

/* what thread 0 needs from a slot to assemble the chain's records (writing pass) */
struct LfSlotRec {
    uint32_t flags;
    uint32_t endc, endm;       /* text offsets (chain-relative) just behind the part that closes a record */
    uint32_t invc, invm;       /* ... and behind the inversion record */
    uint32_t startc, startm;   /* ... where the record opened by a split starts */
    int32_t ed_pre, ed_inv, ed_post;
    uint32_t posEnd, qEnd, newpos, newqs;
};

/* exclusive prefix sum over the block; `total` = sum over all threads.  ws: LF_EMIT_BLOCK / 32 + 1 words. */
__device__ __forceinline__ uint32_t lf_block_excl_scan(uint32_t v, uint32_t *ws, uint32_t &total)
{
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(LF_FULL, incl, o); if (lane >= (uint32_t)o) incl += u; }
    __syncthreads();               /* ws may still be read from the previous scan */
    if (lane == 31u) ws[wid] = incl;
    __syncthreads();
    uint32_t base = 0; total = 0;
#pragma unroll
    for (int w = 0; w < LF_EMIT_BLOCK / 32; w++) { const uint32_t t = ws[w]; if ((uint32_t)w < wid) base += t; total += t; }
    return base + incl - v;
}

/* block-wide copy of len bytes from shared memory to global memory; src and dst agree modulo 16 */
__device__ __forceinline__ void lf_stage_copy_out(char *dst, const char *src, uint32_t len)
{
    const uint32_t tid = threadIdx.x;
    uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > len) head = len;
    if (tid < head) dst[tid] = src[tid];
    const uint32_t body = (len - head) >> 4;
    const uint4 *s4 = (const uint4 *)(src + head);
    uint4 *d4 = (uint4 *)(dst + head);
    for (uint32_t x = tid; x < body; x += LF_EMIT_BLOCK) d4[x] = s4[x];
    const uint32_t done = head + (body << 4);
    if (tid < len - done) dst[done + tid] = src[done + tid];
}

template <bool WRITE>
__global__ void __launch_bounds__(LF_EMIT_BLOCK) k_emit_slots(LfEmitDev d)
{
    __shared__ uint32_t s_ws[LF_EMIT_BLOCK / 32 + 1];
    __shared__ uint32_t s_tail_c[LF_EMIT_BLOCK], s_tail_m[LF_EMIT_BLOCK];
    __shared__ uint8_t s_pass[LF_EMIT_BLOCK];
    __shared__ LfSlotRec s_rec[WRITE ? LF_EMIT_BLOCK : 1];
    __shared__ uint4 s_stage_c4[WRITE ? LF_EMIT_STAGE / 16 : 1], s_stage_m4[WRITE ? LF_EMIT_STAGE / 16 : 1];
    char *const s_stage_c = (char *)s_stage_c4, *const s_stage_m = (char *)s_stage_m4;
    __shared__ uint32_t s_nrec;
    const uint32_t li = blockIdx.x, tid = threadIdx.x;
    const uint32_t c = d.chain_list ? d.chain_list[li] : li;
    const lf_chain ch = d.chains[c];
    const lf_seed *s = d.seeds + ch.seed_off;
    const uint32_t n = ch.n_seeds, nslots = n + 1;
    const uint32_t readLen = (uint32_t)(d.read_off[ch.read_id + 1] - d.read_off[ch.read_id]);
    const uint32_t flag_norm = ch.is_rev ? 16u : 0u, flag_opp = ch.is_rev ? 0u : 16u;
    const uint8_t guards = d.guards[c];
    int32_t clip[4] = { -1, -1, -1, -1 };
    if (d.clip) { clip[0] = d.clip[4 * (size_t)c]; clip[1] = d.clip[4 * (size_t)c + 1]; clip[2] = d.clip[4 * (size_t)c + 2]; clip[3] = d.clip[4 * (size_t)c + 3]; }
    const uint32_t sp_begin = d.split_begin ? d.split_begin[c] : 0u, sp_end = d.split_begin ? d.split_begin[c + 1] : 0u;
    const uint64_t task0 = d.task_base[c], slot0 = d.slot_base[c];
    /* does the chain's first record die at gap 0 (fewer than two anchors)? */
    bool head_dropped = false;
    for (uint32_t sp = sp_begin; sp < sp_end; sp++) if (d.splits[sp].split) { head_dropped = d.splits[sp].gap_i == 0; break; }

    /* running state across rounds of LF_EMIT_BLOCK slots (uniform over the block) */
    uint32_t run_tasks = 0, run_c = 0, run_m = 0, run_carry_c = 0, run_carry_m = 0;
    if (!WRITE && tid == 0) s_nrec = 0;
    /* thread 0 only (writing pass) */
    uint32_t cur_startc = 0, cur_startm = 0, cur_pos = 0, cur_qs = 0, rec_n = 0;
    int32_t acc_ed = 0;
    uint64_t cig_base = 0, md_base = 0, rec_base = 0;
    if (WRITE) { cig_base = d.cig_off[li]; md_base = d.cig_off[d.n_chains] + d.md_off[li]; rec_base = d.rec_off[li]; }

    for (uint32_t base = 0; base < nslots; base += LF_EMIT_BLOCK) {
        const uint32_t k = base + tid;
        const bool active = k < nslots;
        /* ---- geometry of the slot ---- */
        lf_seed s0; s0.tPos = 0; s0.qPos = 0; s0.len = 0;
        lf_seed s1 = s0;
        int32_t ql = 0, tl = 0;
        bool has_task = false;
        if (active) {
            if (k == 0) { s1 = s[0]; has_task = s1.qPos > 0 && (guards & 1); }
            else if (k < n) { s0 = s[k - 1]; s1 = s[k]; ql = (int32_t)(s1.qPos - (s0.qPos + s0.len)); tl = (int32_t)(s1.tPos - (s0.tPos + s0.len)); has_task = ql > 0 && tl > 0; }
            else { s0 = s[n - 1]; has_task = (int32_t)readLen - (int32_t)(s0.qPos + s0.len) > 0 && (guards & 2); }
        }
        uint32_t trel, carry_c = 0, carry_m = 0, off_c = 0, off_m = 0;
        if (!WRITE) {
            uint32_t tot;
            trel = run_tasks + lf_block_excl_scan(has_task ? 1u : 0u, s_ws, tot);
            run_tasks += tot;
        } else {
            trel = 0;
            if (active) { const LfSlotInfo si = d.slot_info[slot0 + k]; carry_c = si.carry_c; carry_m = si.carry_m; off_c = si.off_c; off_m = si.off_m; trel = d.slot_task[slot0 + k]; }
        }
        /* ---- where this round's text goes: staged in shared memory and copied out in 16-byte pieces when it
         *      fits, else (long deletions, inversion records) written in place byte by byte ---- */
        uint32_t rnd_c0 = 0, rnd_c1 = 0, rnd_m0 = 0, rnd_m1 = 0;   /* chain-relative byte span of the round */
        bool staged = false;
        if (WRITE) {
            const LfSlotInfo f = d.slot_info[slot0 + base];
            rnd_c0 = f.off_c; rnd_m0 = f.off_m;
            if (base + LF_EMIT_BLOCK < nslots) { const LfSlotInfo g = d.slot_info[slot0 + base + LF_EMIT_BLOCK]; rnd_c1 = g.off_c; rnd_m1 = g.off_m; }
            else { rnd_c1 = (uint32_t)(d.cig_off[li + 1] - d.cig_off[li]); rnd_m1 = (uint32_t)(d.md_off[li + 1] - d.md_off[li]); }
            staged = rnd_c1 - rnd_c0 + 16u <= LF_EMIT_STAGE && rnd_m1 - rnd_m0 + 16u <= LF_EMIT_STAGE;
        }
        /* a staged byte sits at the same address modulo 16 as its destination, so the copy-out is aligned */
        const uint32_t sk_c = WRITE ? (uint32_t)((uintptr_t)(d.text + cig_base + rnd_c0) & 15u) : 0u, sk_m = WRITE ? (uint32_t)((uintptr_t)(d.text + md_base + rnd_m0) & 15u) : 0u;
        /* ---- walk: every part of a slot is  run(c1, n1) . [alignment ops | deleted bases] . run(c3, n3) ---- */
        LfSlotB<WRITE> B;
        B.cn = 0; B.mn = 0; B.cp = nullptr; B.mp = nullptr;
        if (WRITE) {
            if (staged) { B.cp = s_stage_c + sk_c + (off_c - rnd_c0); B.mp = s_stage_m + sk_m + (off_m - rnd_m0); }
            else { B.cp = d.text + cig_base + off_c; B.mp = d.text + md_base + off_m; }
        }
        uint32_t flags = 0;
        uint32_t pre_c = 0, pre_m = 0, inv_c = 0, inv_m = 0;     /* bytes behind the pre part / the inversion record */
        uint32_t cfirst = 0, mfirst = 0; bool cclosed = false, mclosed = false;
        uint32_t tail_c = 0, tail_m = 0; bool pass_c = false, pass_m = false;
        int32_t ed_pre = 0, ed_inv = 0, ed_post = 0;
        uint32_t posEnd = 0, qEnd = 0, newpos = 0, newqs = 0;
        uint32_t my_nrec = 0;
        if (active) {
            const uint64_t task = task0 + trel;
            /* part descriptors: [0] belongs to the record the slot starts in, [1] to the record a split opens */
            bool on[2] = { false, false }, fresh[2] = { false, true }, fin[2] = { false, false }, rev[2] = { false, false };
            char c1[2] = { 'M', 'I' }, c3[2] = { 'I', 'I' };
            int32_t n1[2] = { 0, 0 }, n3[2] = { 0, 0 }, seg[2] = { 0, 0 }, dtl = 0;   /* seg: 0 none, 1 round-1 ops, 3 round-3 ops, 2 deleted bases */
            int32_t rix[2] = { 0, 0 };
            uint32_t t0[2] = { 0, 0 };
            const LfSplitDev *sv = nullptr;
            if (k == 0) {                                   /* head (:1820-1899) */
                flags |= LF_SL_HEAD;
                newpos = s1.tPos; newqs = s1.qPos;
                const int32_t a = (int32_t)s1.qPos;
                fresh[0] = true; c1[0] = 'I';
                if (a > 0 && !head_dropped) {
                    on[0] = true;
                    if (guards & 1) {
                        rev[0] = true;
                        if (clip[0] >= 0) { n1[0] = a - clip[1]; seg[0] = 3; rix[0] = clip[0]; newqs = s1.qPos - (uint32_t)clip[1]; }
                        else { seg[0] = 1; newqs = 0; }
                    } else n1[0] = a;
                }
            } else if (k < n) {                             /* anchor k-1 and the gap behind it (:1901-2137) */
                const uint32_t i = k - 1;
                const uint32_t ts = s0.tPos + s0.len;
                uint32_t svi = 0;
                if (has_task) for (uint32_t sp = sp_begin; sp < sp_end; sp++) if (d.splits[sp].gap_i == i) { sv = &d.splits[sp]; svi = sp; break; }
                if (sv && !sv->split) sv = nullptr;
                n1[0] = (int32_t)s0.len; t0[0] = ts;
                if (sv) {
                    flags |= LF_SL_SPLIT;
                    /* anchors in the record this gap closes, and in the one it opens: fewer than two -> dropped */
                    bool drop_pre = i == 0, drop_post = false;
                    for (uint32_t sp = svi; sp > sp_begin; sp--) if (d.splits[sp - 1].split) { drop_pre = i - d.splits[sp - 1].gap_i <= 1u; break; }
                    for (uint32_t sp = svi + 1; sp < sp_end; sp++) if (d.splits[sp].split) { drop_post = d.splits[sp].gap_i == i + 1u; break; }
                    posEnd = sv->ts2; qEnd = sv->qs2; newpos = sv->te2; newqs = sv->qe2;
                    on[0] = !drop_pre; fin[0] = true;
                    if (sv->t_first >= 0) { seg[0] = 3; rix[0] = sv->t_first; }
                    n3[0] = (int32_t)(readLen - sv->qs2);
                    on[1] = !drop_post;
                    n1[1] = (int32_t)sv->qe2;
                    if (sv->t_second >= 0) { seg[1] = 3; rix[1] = sv->t_second; rev[1] = true; t0[1] = sv->te2; }
                } else {
                    on[0] = true;
                    if (has_task) seg[0] = 1;
                    else if (ql > 0) { n3[0] = ql; ed_pre -= ql; }
                    else { seg[0] = 2; dtl = tl; ed_pre -= tl; }
                }
            } else {                                        /* last anchor and the tail (:2139-2230) */
                flags |= LF_SL_FINAL;
                on[0] = true; fin[0] = true;
                n1[0] = (int32_t)s0.len;
                posEnd = s0.tPos + s0.len - 1;
                qEnd = s0.qPos + s0.len - 1;                /* inclusive, :2155 */
                const uint32_t qs = s0.qPos + s0.len;
                const int32_t b = (int32_t)readLen - (int32_t)qs;
                t0[0] = s0.tPos + s0.len;
                if (b > 0) {
                    if (guards & 2) {
                        if (clip[2] >= 0) { seg[0] = 3; rix[0] = clip[2]; qEnd = qs + (uint32_t)clip[3]; n3[0] = b - clip[3]; }
                        else { seg[0] = 1; qEnd = readLen; }
                    } else n3[0] = b;
                }
            }
#pragma unroll 1
            for (int p = 0; p < 2; p++) {
                if (p == 1) {
                    pre_c = B.cn; pre_m = B.mn;
                    if (sv && sv->inv) {                    /* the inversion record sits between the two parts */
                        const lf_align_result rr = d.r3[sv->t_mid_r];
                        B.start(true, 0, 0);
                        lf_emit_inversion(B, d.ops3, rr, d.pac, sv->ts2, sv->qs2, readLen - sv->qe2);
                        B.finish();
                        ed_inv = -rr.edit_distance;
                        flags |= LF_SL_INV; my_nrec++;
                    }
                    inv_c = B.cn; inv_m = B.mn;
                }
                if (!on[p]) continue;
                B.start(fresh[p], carry_c, carry_m);
                B.run(c1[p], n1[p]);
                if (seg[p] == 2) {
                    B.cig_run('D', dtl);
                    for (int32_t x = 0; x < dtl; x++) B.md_del(lf_pac_char(d.pac, t0[p] + (uint32_t)x));
                } else if (seg[p]) {
                    const lf_align_result r = seg[p] == 1 ? d.r1[task] : d.r3[rix[p]];
                    const int32_t e = r.edit_distance;
                    if (p == 0) ed_pre -= e; else ed_post -= e;
                    uint32_t tt = t0[p];
                    if (k == 0) { tt = s1.tPos - (uint32_t)(r.end_location + 1); newpos = tt; }
                    else if (k == n) posEnd = tt + (uint32_t)r.end_location;
                    lf_emit_segment(B, seg[p] == 1 ? d.ops1 : d.ops3, r, rev[p], d.pac, tt);
                }
                B.run(c3[p], n3[p]);
                if (fin[p]) { B.finish(); my_nrec++; if (k < n) flags |= LF_SL_PUSH; }
                else B.end_slot();
                if (p == 0 && !fresh[0]) {
                    cfirst = B.cfirst; mfirst = B.mfirst;
                    cclosed = !WRITE && !B.cdef; mclosed = !WRITE && !B.mdef;
                    pass_c = !WRITE && B.cdef; pass_m = !WRITE && B.mdef;
                }
                if (!fin[p]) { tail_c = B.cch == 'M' ? B.cnum : 0u; tail_m = B.mnum; }
            }
        }
        if (!WRITE) {
            /* ---- carries: the open 'M' run / match count at the end of the previous slot ---- */
            s_tail_c[tid] = tail_c; s_tail_m[tid] = tail_m;
            s_pass[tid] = (uint8_t)((pass_c ? 1 : 0) | (pass_m ? 2 : 0));
            __syncthreads();
            if (active) {
                uint32_t acc = 0; int j = (int)tid - 1;
                for (; j >= 0; j--) { acc += s_tail_c[j]; if (!(s_pass[j] & 1)) break; }
                carry_c = j < 0 ? acc + run_carry_c : acc;
                acc = 0; j = (int)tid - 1;
                for (; j >= 0; j--) { acc += s_tail_m[j]; if (!(s_pass[j] & 2)) break; }
                carry_m = j < 0 ? acc + run_carry_m : acc;
            }
            /* the round hands its last slot's state on */
            const uint32_t lastt = (nslots - base < LF_EMIT_BLOCK ? nslots - base : LF_EMIT_BLOCK) - 1;
            const uint32_t nxt_c = pass_c ? carry_c + tail_c : tail_c, nxt_m = pass_m ? carry_m + tail_m : tail_m;
            __syncthreads();
            if (tid == lastt) { s_tail_c[0] = nxt_c; s_tail_m[0] = nxt_m; }
            __syncthreads();
            run_carry_c = s_tail_c[0]; run_carry_m = s_tail_m[0];
            /* ---- sizes and offsets ---- */
            uint32_t bytes_c = 0, bytes_m = 0;
            if (active) {
                const uint32_t post_c = B.cn - inv_c, post_m = B.mn - inv_m;
                bytes_c = pre_c + (cclosed ? lf_ndigits(carry_c + cfirst) : 0u) + (inv_c - pre_c) + post_c;
                bytes_m = pre_m + (mclosed ? lf_ndigits(carry_m + mfirst) : 0u) + (inv_m - pre_m) + post_m;
            }
            uint32_t tot_c, tot_m;
            const uint32_t oc = run_c + lf_block_excl_scan(bytes_c, s_ws, tot_c);
            const uint32_t om = run_m + lf_block_excl_scan(bytes_m, s_ws, tot_m);
            run_c += tot_c; run_m += tot_m;
            if (active) {
                LfSlotInfo si; si.carry_c = carry_c; si.carry_m = carry_m; si.off_c = oc; si.off_m = om;
                d.slot_info[slot0 + k] = si; d.slot_task[slot0 + k] = trel;
                if (my_nrec) atomicAdd(&s_nrec, my_nrec);
            }
            __syncthreads();
        } else {
            if (staged) {   /* ---- staged text -> HBM, 16 bytes per thread and step ---- */
                __syncthreads();
                lf_stage_copy_out(d.text + cig_base + rnd_c0, s_stage_c + sk_c, rnd_c1 - rnd_c0);
                lf_stage_copy_out(d.text + md_base + rnd_m0, s_stage_m + sk_m, rnd_m1 - rnd_m0);
            }
            /* ---- records: thread 0 walks the round's slots in order ---- */
            LfSlotRec R;
            R.flags = active ? flags : 0u;
            R.endc = off_c + pre_c; R.endm = off_m + pre_m;
            R.invc = off_c + inv_c; R.invm = off_m + inv_m;
            R.startc = R.invc; R.startm = R.invm;
            R.ed_pre = ed_pre; R.ed_inv = ed_inv; R.ed_post = ed_post;
            R.posEnd = posEnd; R.qEnd = qEnd; R.newpos = newpos; R.newqs = newqs;
            s_rec[tid] = R;
            __syncthreads();
            if (tid == 0) {
                const uint32_t cnt = nslots - base < LF_EMIT_BLOCK ? nslots - base : LF_EMIT_BLOCK;
                for (uint32_t x = 0; x < cnt; x++) {
                    const LfSlotRec &Q = s_rec[x];
                    if (Q.flags & LF_SL_HEAD) { cur_pos = Q.newpos; cur_qs = Q.newqs; acc_ed = Q.ed_pre; continue; }
                    acc_ed += Q.ed_pre;
                    if (Q.flags & (LF_SL_PUSH | LF_SL_FINAL)) {
                        lf_sam_record rr; rr.chain_id = c; rr.flag = flag_norm; rr.pos = cur_pos; rr.posEnd = Q.posEnd; rr.qStart = cur_qs; rr.qEnd = Q.qEnd; rr.nmCount = acc_ed;
                        rr.cigar_off = d.out_base + cig_base + cur_startc; rr.cigar_len = Q.endc - cur_startc - 1u;
                        rr.md_off = d.out_base + md_base + cur_startm; rr.md_len = Q.endm - cur_startm - 1u;
                        d.recs[rec_base + rec_n++] = rr;
                    }
                    if (Q.flags & LF_SL_INV) {
                        lf_sam_record rr; rr.chain_id = c; rr.flag = flag_opp; rr.pos = Q.posEnd; rr.posEnd = Q.newpos; rr.qStart = Q.qEnd; rr.qEnd = Q.newqs; rr.nmCount = Q.ed_inv;
                        rr.cigar_off = d.out_base + cig_base + Q.endc; rr.cigar_len = Q.invc - Q.endc - 1u;
                        rr.md_off = d.out_base + md_base + Q.endm; rr.md_len = Q.invm - Q.endm - 1u;
                        d.recs[rec_base + rec_n++] = rr;
                    }
                    if (Q.flags & LF_SL_SPLIT) { cur_startc = Q.startc; cur_startm = Q.startm; cur_pos = Q.newpos; cur_qs = Q.newqs; acc_ed = Q.ed_post; }
                }
            }
            __syncthreads();
        }
    }
    if (!WRITE && tid == 0) { d.nrec[li] = s_nrec; d.cig_bytes[li] = run_c; d.md_bytes[li] = run_m; }
}

/* The clip / split trigger tests of alignChain_edlib on the round-1 results (src/LordFAST.cpp:1840, :1952,
 * :2175), in the reference's own arithmetic: the similarity is computed in float and compared with a double
 * constant.  Head and tail tasks are the prefix-mode ones.  Triggered task indices are appended to `list`
 * (unordered; the host sorts the few thousand entries). */
/* Round-1 tasks of a chunk of chains, written where k_align_prep expects them (alignChain_edlib's own calls: head SHW
 * :1827-1833, one NW per gap with query and target bases :1936-1941, tail SHW :2164-2168).  Heads and tails longer than
 * _pf_clipLen are asked for distance and end only: their path is used only if the clip test fails or the extension
 * does not shorten them (:1840-1878, :2175-2209), and then round 3 computes it.  One warp per chain; the
 * host has only counted (task_base) and decided the chromosome-boundary guards.  59 MB of tasks per config-2 chunk
 * neither get built on host threads nor cross PCIe. */
__global__ void __launch_bounds__(128) k_chain_tasks(const lf_chain *__restrict__ chains, const lf_seed *__restrict__ seeds, const uint64_t *__restrict__ read_off,
                                                     const uint64_t *__restrict__ task_base, const uint8_t *__restrict__ guards, uint32_t n_chains, lf_align_task *out)
{
    const uint32_t c = blockIdx.x * 4u + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (c >= n_chains) return;
    const lf_chain ch = chains[c];
    const lf_seed *s = seeds + ch.seed_off;
    const uint32_t n = ch.n_seeds;
    const unsigned strand = ch.is_rev ? LF_F_READ_REV : 0;
    const uint32_t g = guards[c];
    uint64_t k = task_base[c];
    lf_align_task t;
    t.read_id = ch.read_id; t.reserved = 0;
    if (g & 1u) {
        if (lane == 0) {
            const uint32_t a = s[0].qPos;
            t.q_off = 0; t.q_len = a; t.t_off = s[0].tPos - (a + 20u); t.t_len = a + 20u; t.flags = (uint16_t)(strand | LF_F_REVERSE_BOTH | (a > LF_CLIP_LEN ? LF_F_NO_PATH : 0)); t.mode = LF_MODE_SHW;
            out[k] = t;
        }
        k++;
    }
    for (uint32_t i0 = 0; i0 + 1 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        bool is = false;
        if (i + 1 < n) {
            const lf_seed s0 = s[i], s1 = s[i + 1];
            const uint32_t qs = s0.qPos + s0.len, ts = s0.tPos + s0.len;
            const int32_t ql = (int32_t)(s1.qPos - qs), tl = (int32_t)(s1.tPos - ts);
            is = ql > 0 && tl > 0;
            t.q_off = qs; t.q_len = (uint32_t)ql; t.t_off = ts; t.t_len = (uint32_t)tl; t.flags = (uint16_t)strand; t.mode = LF_MODE_NW;
        }
        const uint32_t bal = __ballot_sync(LF_FULL, is);
        if (is) out[k + (uint64_t)__popc(bal & ((1u << lane) - 1u))] = t;
        k += (uint64_t)__popc(bal);
    }
    if ((g & 2u) && lane == 0) {
        const lf_seed sl = s[n - 1];
        const uint32_t qs = sl.qPos + sl.len;
        const uint32_t b = (uint32_t)(read_off[ch.read_id + 1] - read_off[ch.read_id]) - qs;
        t.q_off = qs; t.q_len = b; t.t_off = sl.tPos + sl.len; t.t_len = b + 20u; t.flags = (uint16_t)(strand | (b > LF_CLIP_LEN ? LF_F_NO_PATH : 0)); t.mode = LF_MODE_SHW;
        out[k] = t;
    }
}

/* Pass A of the chain operator on the device (round 2; it was a loop over every seed of the chunk on the host threads, 9 ms
 * of a call on the 4 cores a rank has on an 8-GPU box): per chain the contig it lies in (bns_pos2rid on the midpoint of
 * the first and last seed, src/BWT.cpp:653-660), the head / tail guards (:1825, :2163), the number of round-1 tasks,
 * whether a task's lengths qualify for a clip / split trigger, and the validation of the chain (seeds inside the read and
 * the reference, in order, without overlap: Chain.cpp:258-266 guarantees it).  One warp per chain.
 * guards[c]: bit 0 head aligned, bit 1 tail aligned, bit 2 trigger candidate. */
__global__ void __launch_bounds__(128) k_chain_plan(const lf_chain *__restrict__ chains, const lf_seed *__restrict__ seeds, const uint64_t *__restrict__ read_off,
                                                    uint32_t n_reads, const int64_t *__restrict__ coff, const int32_t *__restrict__ clen, int nctg, int64_t l_pac,
                                                    uint32_t n_chains, uint8_t *guards, uint32_t *ntask, uint32_t *bad)
{
    const uint32_t c = blockIdx.x * 4u + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (c >= n_chains) return;
    const lf_chain ch = chains[c];
    const uint32_t n = ch.n_seeds;
    if (n < 2u || ch.read_id >= n_reads) { if (lane == 0) { atomicOr(bad, 1u); guards[c] = 0; ntask[c] = 0; } return; }
    const lf_seed *s = seeds + ch.seed_off;
    const uint32_t readLen = (uint32_t)(read_off[ch.read_id + 1] - read_off[ch.read_id]);
    uint32_t cnt = 0;
    bool cand = false, good = true;
    for (uint32_t i = lane; i < n; i += 32u) {
        const lf_seed a = s[i];
        good = good && a.len >= 1u && (uint64_t)a.qPos + a.len <= readLen && (int64_t)a.tPos + a.len <= l_pac;
        if (i + 1u < n) {
            const lf_seed b = s[i + 1];
            good = good && b.qPos >= a.qPos + a.len && b.tPos >= a.tPos + a.len;
            const int32_t ql = (int32_t)(b.qPos - (a.qPos + a.len)), tl = (int32_t)(b.tPos - (a.tPos + a.len));
            if (ql > 0 && tl > 0) { cnt++; cand = cand || ql - tl >= 80 || tl - ql >= 80; }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(LF_FULL, cnt, o);
    cand = __any_sync(LF_FULL, cand);
    good = __all_sync(LF_FULL, good);
    if (lane == 0) {
        const lf_seed s0 = s[0], sl = s[n - 1];
        const int64_t pos = ((int64_t)s0.tPos + (int64_t)sl.tPos) >> 1;
        int left = 0, mid = 0, right = nctg;   /* bns_pos2rid, lib/bwa/bntseq.c:349-363 */
        bool found = pos < l_pac;
        while (found && left < right) {
            mid = (left + right) >> 1;
            if (pos >= coff[mid]) {
                if (mid == nctg - 1) break;
                if (pos < coff[mid + 1]) break;
                left = mid + 1;
            } else right = mid;
        }
        uint32_t g = 0;
        if (found && good) {
            const int64_t chrBeg = coff[mid], chrEnd = coff[mid] + clen[mid] - 1;
            const int32_t a = (int32_t)s0.qPos;
            const int32_t b = (int32_t)readLen - (int32_t)(sl.qPos + sl.len);
            const bool hg = a > 0 && (int64_t)s0.tPos - (a + 20) >= (int64_t)(uint32_t)chrBeg;
            const bool tg = b > 0 && sl.tPos + sl.len + (uint32_t)(b + 20) - 1u <= (uint32_t)chrEnd;
            cand = cand || (hg && a > LF_CLIP_LEN) || (tg && b > LF_CLIP_LEN);
            g = (hg ? 1u : 0u) | (tg ? 2u : 0u) | (cand ? 4u : 0u);
            cnt += (hg ? 1u : 0u) + (tg ? 1u : 0u);
        } else { atomicOr(bad, 1u); cnt = 0; }
        guards[c] = (uint8_t)g;
        ntask[c] = cnt;
    }
}

__global__ void k_chain_triggers(const lf_align_task *tasks, const lf_align_result *res, uint32_t n, uint32_t *list, uint32_t *count, uint32_t cap)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const lf_align_task t = tasks[i];
    const int32_t ed = res[i].edit_distance;
    const int32_t ql = (int32_t)t.q_len, tl = (int32_t)t.t_len;
    const double sim = (double)(1 - ((float)ed / ql));
    bool trig;
    if (t.mode == LF_MODE_SHW) trig = ql > 500 && sim < 0.75;
    else { const int32_t dl = ql > tl ? ql - tl : tl - ql; trig = dl >= 80 && sim < 0.40; }
    if (trig) { const uint32_t p = atomicAdd(count, 1u); if (p < cap) list[p] = i; }
}

/* ------------------------------------------------------------------------------------------ */
/* k_myers_band: register-resident Myers with a sliding word band and op planes kept in HBM      */
/* ------------------------------------------------------------------------------------------ */
/* Second-generation small-task kernel.  Two changes against k_myers_small:
 *  (1) The traceback no longer recomputes: the forward pass writes the two op planes of every column to
 *      HBM in a warp-interleaved layout ([column][window word][lane], 256 contiguous bytes per warp store),
 *      and the walk loads its 2-word window for 8 columns at a time.  (Recompute was 42 % of the executed
 *      instructions; the stores are asynchronous and coalesced.)
 *  (2) BANDED: only NB consecutive words (a band of 32*NB rows that slides down the diagonal one word at a
 *      time) are computed and stored, as edlib does with its block band.  Cells outside the band are treated
 *      as "+1 per step" (upper bounds).  With x = 16*(NB-1)-4 rows guaranteed between the straight line
 *      (0,0)-(q,t) and the band edges, a path leaving the band costs >= 2*(x+1) - |q-t| edits, so a computed
 *      distance d <= 32*(NB-1) - 7 - |q-t| proves that no optimal path leaves the band: d, and every
 *      Up/Left/Diagonal test on the path, are then exact (the results are
 *      band-independent).  Tasks that fail the test are flagged LF_RETRY and redone by the full-width
 *      k_myers_small. */
#define LF_RETRY 1
#ifdef LF_EMU
static unsigned long lf_emu_band_ok = 0, lf_emu_band_retry = 0; /* test-only visibility into the certificate */
#define LF_BAND_COUNT(x) ((x)++)
#else
#define LF_BAND_COUNT(x) ((void)0)
#endif
#define LF_BAND_C 8

__host__ __device__ __forceinline__ uint32_t lf_bucket_hi(uint32_t t)
{ /* largest target length in t's sort bucket (see k_align_prep) */
    constexpr uint32_t MB = LF_BUCKET_BITS;
    if (t < (1u << MB)) return t ? t : 1u;
    uint32_t e = 31u;
    while (!(t >> e)) e--;
    const uint32_t m = (t >> (e - MB)) & ((1u << MB) - 1u);
    return (((1u << MB) + m + 1u) << (e - MB)) - 1u;
}

struct LfGroupCfg { uint32_t first[LF_CLS_LARGE], count[LF_CLS_LARGE], gbase[LF_CLS_LARGE + 1], nb[LF_CLS_LARGE]; };

__global__ void k_group_scratch(LfDev d, const uint32_t *__restrict__ order, LfGroupCfg cfg, uint32_t *gbytes)
{ /* one thread per warp group (32 consecutive sorted tasks of one class): bytes of its plane region */
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= cfg.gbase[LF_CLS_LARGE]) return;
    int cls = 0;
    while (cls + 1 < LF_CLS_LARGE && g >= cfg.gbase[cls + 1]) cls++;
    const uint32_t wg = g - cfg.gbase[cls];
    uint32_t bytes = 0;
    if (cfg.nb[cls]) {
        const uint32_t t = d.tasks[order[cfg.first[cls] + wg * 32u]].t_len; /* sorted: the group's first task has its largest bucket */
        bytes = lf_bucket_hi(t) * cfg.nb[cls] * 256u;
    }
    gbytes[g] = bytes;
}

template <int NB, bool SHW>
__device__ __forceinline__ void lf_band_column(uint32_t (&Pv)[NB], uint32_t (&Mv)[NB], const uint32_t (&qlo)[NB], const uint32_t (&qhi)[NB],
                                               const uint32_t (&qnn)[NB], uint32_t slo, uint32_t shi, int &score, int wl_rel, uint32_t bl, uint2 *dst)
{
    uint32_t Eq[NB], a[NB], sum[NB];
#pragma unroll
    for (int w = 0; w < NB; w++) {
        Eq[w] = ~((qlo[w] ^ slo) | (qhi[w] ^ shi) | qnn[w]);
        a[w] = Eq[w] & Pv[w];
    }
    lf_add_chain<NB>(a, Pv, sum);
    uint32_t pPh = 0x80000000u, pMh = 0u; /* the row above the band grows by one per column (true for row 0, an upper bound otherwise) */
#pragma unroll
    for (int w = 0; w < NB; w++) {
        const uint32_t Xh = (sum[w] ^ Pv[w]) | Eq[w];
        const uint32_t Ph = Mv[w] | ~(Xh | Pv[w]);
        const uint32_t Mh = Pv[w] & Xh;
        const uint32_t Xv = Eq[w] | Mv[w];
        if (SHW) { if (w == wl_rel) score += (int)((Ph >> bl) & 1u) - (int)((Mh >> bl) & 1u); }
        const uint32_t Phs = __funnelshift_l(pPh, Ph, 1), Mhs = __funnelshift_l(pMh, Mh, 1);
        pPh = Ph; pMh = Mh;
        const uint32_t nPv = Mhs | ~(Xv | Phs);
        const uint32_t nMv = Phs & Xv;
        const uint32_t diagx = ~(nPv | Ph | Eq[w]);
        dst[w * 32] = make_uint2(nPv | diagx, (~nPv & Ph) | diagx);
        Pv[w] = nPv; Mv[w] = nMv;
    }
}

template <int NB, bool BANDED, bool SHW>
__global__ void __launch_bounds__(128) k_myers_band(LfDev d, const uint32_t *__restrict__ order, uint32_t first, uint32_t count, uint32_t gbase,
                                                    const unsigned long long *__restrict__ goff, uint32_t *retry_list, uint32_t *retry_count)
{
    static_assert(!(BANDED && SHW), "prefix-mode tasks run unbanded");
    constexpr int WIN = NB < 2 ? 1 : 2;
    constexpr int C = LF_BAND_C;
    LF_DYN_SMEM(uint32_t, smem); /* [C][WIN][2][128] */
    const uint32_t tid = threadIdx.x;
    const uint32_t gi = blockIdx.x * 128u + tid;
    if (gi >= count) return;
    const uint32_t ti = order[first + gi];
    const lf_align_task task = d.tasks[ti];
    const int q = (int)task.q_len, t = (int)task.t_len;
    const int nw = (q + 31) >> 5;
    lf_align_result r;
    r.status = 0; r.ops_len = 0;
    const uint64_t slot_hi = d.slot_end[ti] * 16ull;
    r.ops_off = slot_hi;
    const int kmax = BANDED ? (nw > NB ? nw - NB : 0) : 0;
    const int dq = q / t, dr = q % t;           /* the line's row advances dq (+1 on carry) per column */
    if (BANDED && kmax > 0) {
        const int dlt = q > t ? q - t : t - q;
        if (dq >= 32 || 32 * (NB - 1) - 7 - dlt < 0) { LF_BAND_COUNT(lf_emu_band_retry); retry_list[first + atomicAdd(retry_count, 1u)] = ti; return; }
    }
    LfQView qv; LfTView tv;
    lf_task_views(d, task, qv, tv);
    uint2 *G = (uint2 *)(d.planes + goff[gbase + (gi >> 5)]) + (tid & 31u); /* element (column c, band word j) at G[(c*NB + j)*32] */

    uint32_t qlo[NB], qhi[NB], qnn[NB], Pv[NB], Mv[NB];
#pragma unroll
    for (int w = 0; w < NB; w++) {
        if (w < nw) lf_q32(d, qv, (int64_t)w * 32, qlo[w], qhi[w], qnn[w]);
        else { qlo[w] = 0; qhi[w] = 0; qnn[w] = 0xffffffffu; }
        Pv[w] = 0xffffffffu; Mv[w] = 0u;
    }
    const int wl = (q - 1) >> 5;
    const uint32_t bl = (uint32_t)(q - 1) & 31u;
    int score = q, best = q, bestc = -1;
    int k = 0;                 /* top word of the band */
    int rline = 0, racc = 0;   /* floor((c+1)*q/t) and its remainder */
    int top = 0;               /* D(32k, c) carried along the band's upper edge */

    LfTCursor tc;
    tc.init(d.pac, tv.t0, tv.dir);
    for (int c = 0; c < t; c++) {
        if (BANDED) {
            rline += dq; racc += dr;
            if (racc >= t) { racc -= t; rline++; }
            int kn = (rline - 16 * NB + 16) >> 5;
            kn = kn < 0 ? 0 : kn > kmax ? kmax : kn;
            if (kn != k) { /* slide one word down: the top word's vertical deltas move into `top` */
                top += __popc(Pv[0]) - __popc(Mv[0]);
#pragma unroll
                for (int w = 0; w + 1 < NB; w++) { Pv[w] = Pv[w + 1]; Mv[w] = Mv[w + 1]; qlo[w] = qlo[w + 1]; qhi[w] = qhi[w + 1]; qnn[w] = qnn[w + 1]; }
                Pv[NB - 1] = 0xffffffffu; Mv[NB - 1] = 0u;
                k++;
                if (k + NB - 1 < nw) lf_q32(d, qv, (int64_t)(k + NB - 1) * 32, qlo[NB - 1], qhi[NB - 1], qnn[NB - 1]);
                else { qlo[NB - 1] = 0; qhi[NB - 1] = 0; qnn[NB - 1] = 0xffffffffu; }
            }
        }
        uint32_t slo, shi;
        tc.next_masks(slo, shi);
        lf_band_column<NB, SHW>(Pv, Mv, qlo, qhi, qnn, slo, shi, score, wl, bl, G + (size_t)c * (NB * 32));
        if (SHW) { if (score < best) { best = score; bestc = c; } }
    }
    int ed, end;
    if (SHW) { ed = best; end = bestc; }
    else {
        ed = t + top;
#pragma unroll
        for (int w = 0; w < NB; w++) {
            const int wa = k + w;
            const uint32_t m = wa < wl ? 0xffffffffu : wa == wl ? (0xffffffffu >> (31u - bl)) : 0u;
            ed += __popc(Pv[w] & m) - __popc(Mv[w] & m);
        }
        end = t - 1;
    }
    if (BANDED && kmax > 0) {
        const int dlt = q > t ? q - t : t - q;
        if (ed > 32 * (NB - 1) - 7 - dlt) { LF_BAND_COUNT(lf_emu_band_retry); retry_list[first + atomicAdd(retry_count, 1u)] = ti; return; }
        LF_BAND_COUNT(lf_emu_band_ok);
    }
    r.edit_distance = ed; r.end_location = end;
    if (task.flags & LF_F_NO_PATH) { d.res[ti] = r; return; }

    /* ---- traceback over the stored planes: Up > Left > Diagonal (edlib.cpp:950, :984, :1015) ---- */
    LfOpSink sink;
    sink.init(slot_hi);
    int i = q, j = end + 1;
    uint32_t *smt = smem + tid;
    constexpr int CS = WIN * 2 * 128;
    bool lost = false;
    /* plane words of one block (C columns x WIN words) travel global -> registers -> shared memory; the registers
     * for the NEXT block are requested before the current block is walked, so the HBM/L2 latency hides behind
     * the walk.  The next block's window is predicted from the current row; a wrong guess only costs a reload. */
    uint2 pre[C * WIN];
    int pre_c0 = -1, pre_wtop = 0;
    auto load_block = [&](int c0, int ncol, int wtop) {
        int kc = 0, rl = 0, ra = 0; /* band position at column c0, then incrementally */
        if (BANDED && kmax > 0) {
            const long long num = (long long)(c0 + 1) * q;
            rl = (int)(num / t); ra = (int)(num % t);
            kc = (rl - 16 * NB + 16) >> 5; kc = kc < 0 ? 0 : kc > kmax ? kmax : kc;
        }
#pragma unroll
        for (int cc = 0; cc < C; cc++) {
#pragma unroll
            for (int wi = 0; wi < WIN; wi++) {
                const int rel = wtop + wi - kc;
                uint2 v = make_uint2(0u, 0u);
                if (cc < ncol && rel >= 0 && rel < NB) v = G[((size_t)(c0 + cc) * NB + rel) * 32];
                pre[cc * WIN + wi] = v;
            }
            if (BANDED && kmax > 0) {
                rl += dq; ra += dr;
                if (ra >= t) { ra -= t; rl++; }
                kc = (rl - 16 * NB + 16) >> 5; kc = kc < 0 ? 0 : kc > kmax ? kmax : kc;
            }
        }
    };
    while (i > 0 && j > 0) {
        const int c1 = j, c0 = ((j - 1) / C) * C;
        const int whi = (i - 1) >> 5;
        int wtop = whi - WIN + 1 > 0 ? whi - WIN + 1 : 0; /* absolute words wtop .. wtop+WIN-1 */
        if (pre_c0 == c0 && c1 == c0 + C && whi >= pre_wtop && whi < pre_wtop + WIN) wtop = pre_wtop; /* the prefetched block fits */
        else load_block(c0, c1 - c0, wtop);
#pragma unroll
        for (int e = 0; e < C * WIN; e++) { smt[e * 2 * 128] = pre[e].x; smt[(e * 2 + 1) * 128] = pre[e].y; }
        if (c0 > 0) { /* request the block to the left now; it is consumed after this block's walk */
            const int ip = i - 1 - C > 0 ? i - 1 - C : 0;
            const int wp = ip >> 5;
            pre_wtop = wp - WIN + 1 > 0 ? wp - WIN + 1 : 0;
            pre_c0 = c0 - C;
            load_block(pre_c0, C, pre_wtop);
        } else pre_c0 = -1;
        const int rowmin = wtop * 32;
        while (i > 0 && j > c0 && (i - 1) >= rowmin) {
            const int wrow = (i - 1) >> 5;
            lf_walk_row<128, CS>(smem, (int)tid + (wrow - wtop) * 2 * 128, wrow, c0, i, j, sink, d.ops);
        }
        if (sink.nops > (uint32_t)(q + t)) { lost = true; break; } /* cannot happen on a certified band */
    }
    while (i > 0 && !lost) { sink.emit(d.ops, 1u); i--; }
    while (j > 0 && !lost) { sink.emit(d.ops, 2u); j--; }
    sink.finish(d.ops);
    if (lost) { retry_list[first + atomicAdd(retry_count, 1u)] = ti; return; }
    const uint32_t nops = sink.nops;
    r.ops_off = slot_hi - nops; r.ops_len = nops;
    d.res[ti] = r;
}

/* ------------------------------------------------------------------------------------------ */
/* k_myers_bandreg: thread per task, NB-word sliding band held in registers                    */
/* ------------------------------------------------------------------------------------------ */
/* Global-mode gap tasks with 128 < q <= 512 hold 54 % of the stage's word-columns (profiles/r02a) and are
 * near-diagonal (|q-t| small, distance ~0.15 q), so only a band of NB words around the line (0,0)-(q,t) can
 * hold an optimal path -- the Ukkonen argument behind edlib's own band (edlib.cpp:722-760).  This kernel is
 * k_myers_small restricted to that band: the band geometry and the certificate are k_myers_band's (see the
 * comment above LF_RETRY: a distance d <= 32*(NB-1) - 7 - |q-t| proves that no optimal path leaves the band,
 * which makes d and every Up/Left/Diagonal test on the path exact), but nothing per column goes to HBM:
 *   forward    NB words per column, (Pv,Mv) of the band checkpointed every 8 columns; blocks of 8 columns in
 *              which the band does not slide (3 of 4 when q ~ t) run fully unrolled: the 8 target symbols are
 *              one funnel shift out of a 64-bit window of the 2-bit reference, strand handled once per word.
 *   traceback  per 8-column block: restore the checkpoint, recompute the band with the two op planes of a
 *              2-word window written to shared memory, walk Up > Left > Diagonal inside the window.
 * Tasks that fail the certificate are appended to the retry list and redone full width by k_myers_small.
 * k_align_prep routes only tasks with 3*|q-t| <= 32*(NB-1)-7 here (classes LF_CLS_BANDREG0..+3). */
template <int NB, bool STORE, int WIN = 2>
__device__ __forceinline__ void lf_bandreg_column(uint32_t (&Pv)[NB], uint32_t (&Mv)[NB], const uint32_t (&qlo)[NB], const uint32_t (&qhi)[NB],
                                                  const uint32_t (&qnn)[NB], uint32_t slo, uint32_t shi, uint32_t *sm, int wrel0)
{ /* STORE: band words wrel0, wrel0+1 are the window words 0, 1 of this column */
    uint32_t nEq[NB], a[NB], sum[NB];
#pragma unroll
    for (int w = 0; w < NB; w++) {
        nEq[w] = lf_neq(qlo[w], qhi[w], qnn[w], slo, shi);
        a[w] = Pv[w] & ~nEq[w];
    }
    lf_add_chain<NB>(a, Pv, sum);
    uint32_t pPh = 0x80000000u, pMh = 0u; /* the row above the band grows by one per column (exact for row 0, an upper bound otherwise) */
#pragma unroll
    for (int w = 0; w < NB; w++) {
        const uint32_t Xh = (sum[w] ^ Pv[w]) | ~nEq[w];
        const uint32_t Ph = Mv[w] | ~(Xh | Pv[w]);
        const uint32_t Mh = Pv[w] & Xh;
        const uint32_t Xv = ~nEq[w] | Mv[w];
        const uint32_t Phs = __funnelshift_l(pPh, Ph, 1), Mhs = __funnelshift_l(pMh, Mh, 1);
        pPh = Ph; pMh = Mh;
        const uint32_t nPv = Mhs | ~(Xv | Phs);
        const uint32_t nMv = Phs & Xv;
        if (STORE) {
            const unsigned wi = (unsigned)(w - wrel0);
            if (wi < (unsigned)WIN) {
                const uint32_t diagx = ~(nPv | Ph) & nEq[w];          /* diagonal step over a mismatch */
                sm[(wi * 2 + 0) * 128] = nPv | diagx;               /* ops 1, 3 */
                sm[(wi * 2 + 1) * 128] = (~nPv & Ph) | diagx;       /* ops 2, 3 */
            }
        }
        Pv[w] = nPv; Mv[w] = nMv;
    }
}

/* Column at whose start the band slides from k to k+1: the first c with floor((c+1)*q/t) - 16*NB + 16 >= 32*(k+1)
 * (the per-column rule of k_myers_band, solved for c: one division per slide instead of a test per column). */
template <int NB>
__device__ __forceinline__ int lf_bandreg_next_slide(int k, int kmax, int q, int t)
{
    if (k >= kmax) return 0x7fffffff;
    const uint32_t R = (uint32_t)(32 * (k + 1) + 16 * NB - 16);
    return (int)((R * (uint32_t)t + (uint32_t)q - 1u) / (uint32_t)q) - 1;
}

/* One block of n <= 8 columns starting at column c with the stream window tb.  A slide (register moves plus the
 * plane words of one more query word) sits behind a test that is uniform for the warp unless one of its tasks
 * slides at this very column.
 * FWD accumulates the vertical deltas that leave through the top of the band (the distance needs them); STORE
 * writes the window planes. */
template <int NBF, int NB, bool FWD, bool STORE, bool FULL, bool SL = true>
__device__ __forceinline__ void lf_bandreg_block(uint32_t (&Pv_)[NBF], uint32_t (&Mv_)[NBF], uint32_t (&qlo_)[NBF], uint32_t (&qhi_)[NBF], uint32_t (&qnn_)[NBF],
                                                 uint32_t tb, int c, int n, int &k, int &cslide, int &top, int q, int t, int kmax,
                                                 const LfDev &d, const LfQView &qv, uint32_t *smt, int wtop)
{ /* NBF: words of the band; NB <= NBF: the top words of it this call computes (the traceback needs none below its row:
   * a word only feeds the words under it, and a word that slides in uncomputed lies below the row as well) */
    static_assert(NB <= NBF, "");
    uint32_t (&Pv)[NB] = reinterpret_cast<uint32_t (&)[NB]>(Pv_);
    uint32_t (&Mv)[NB] = reinterpret_cast<uint32_t (&)[NB]>(Mv_);
    uint32_t (&qlo)[NB] = reinterpret_cast<uint32_t (&)[NB]>(qlo_);
    uint32_t (&qhi)[NB] = reinterpret_cast<uint32_t (&)[NB]>(qhi_);
    uint32_t (&qnn)[NB] = reinterpret_cast<uint32_t (&)[NB]>(qnn_);
    constexpr int WIN = NBF < 2 ? 1 : 2;
    constexpr int CS = WIN * 2 * 128;
    const int es = cslide - c;                         /* the block slides before its column es, if 0 <= es < n */
    const int nn = FULL ? 8 : n;
    /* One copy of the column body (unrolled by two) and one of the slide: ~20 size-class kernels share an SM's
     * instruction cache when the classes run concurrently, and fully unrolled 8-column bodies (13-50 KB per kernel)
     * made all of them stall on instruction fetches. */
    int e = 0;
    int stop = (SL && (unsigned)es < (unsigned)nn) ? es : nn;   /* !SL: the band is the whole column, it never slides */
    for (;;) {
#pragma unroll 2
        for (; e < stop; e++) {
            const uint32_t shi = (uint32_t)((int32_t)tb >> 31), slo = (uint32_t)((int32_t)(tb << 1) >> 31);
            tb <<= 2;
            lf_bandreg_column<NB, STORE, WIN>(Pv, Mv, qlo, qhi, qnn, slo, shi, STORE ? smt + e * CS : nullptr, wtop - k);
        }
        if (e >= nn) break;
        /* slide one word down: the top word's vertical deltas move into `top` */
        if (FWD) top += __popc(Pv[0]) - __popc(Mv[0]);
#pragma unroll
        for (int w = 0; w + 1 < NB; w++) { Pv[w] = Pv[w + 1]; Mv[w] = Mv[w + 1]; }
        Pv[NB - 1] = 0xffffffffu; Mv[NB - 1] = 0u;
#pragma unroll
        for (int w = 0; w + 1 < NBF; w++) { qlo_[w] = qlo_[w + 1]; qhi_[w] = qhi_[w + 1]; qnn_[w] = qnn_[w + 1]; }   /* the query words of the whole band stay in step */
        k++;
        lf_q32(d, qv, (int64_t)(k + NBF - 1) * 32, qlo_[NBF - 1], qhi_[NBF - 1], qnn_[NBF - 1]);   /* k <= kmax: word k+NBF-1 <= nw-1 */
        stop = nn;
    }
    /* slides are more than 8 columns apart (q < 4t: 32 rows of the line take more than 8 columns), so a block holds
     * at most one and the next one is looked up once, here */
    if (SL && (unsigned)es < (unsigned)n) cslide = lf_bandreg_next_slide<NBF>(k, kmax, q, t);
}

template <int NB, bool SLIDE>
__global__ void __launch_bounds__(128) k_myers_bandreg(LfDev d, const uint32_t *__restrict__ order, uint32_t first, uint32_t count,
                                                       uint32_t *retry_list, uint32_t *retry_count, int nwmax)
{
    constexpr int C = 8;
    constexpr int WIN = NB < 2 ? 1 : 2;   /* one-word tasks need one window word: half the shared memory, twice the resident blocks */
    constexpr int CS = WIN * 2 * 128; /* shared-memory words per column: [window words][2 planes][128 threads] */
    LF_DYN_SMEM(uint32_t, smem);    /* window planes [C][2][2][128] */
    const uint32_t tid = threadIdx.x;
    const uint32_t gi = blockIdx.x * 128u + tid;
    if (gi >= count) return;
    const uint32_t ti = order[first + gi];
    const lf_align_task task = d.tasks[ti];
    const int q = (int)task.q_len, t = (int)task.t_len;
    const int nw = (q + 31) >> 5;
    const int kmax = SLIDE ? (nw > NB ? nw - NB : 0) : 0;   /* 0: the band is the whole column and the result needs no certificate (!SLIDE: known at compile time, the slide code goes away) */
    const int dlt = q > t ? q - t : t - q;
    const int cert = 32 * (NB - 1) - 7 - dlt;
    if (nw > nwmax || (!SLIDE && nw > NB) || (kmax > 0 && (q >= 4 * t || cert < 0))) { LF_BAND_COUNT(lf_emu_band_retry); retry_list[first + atomicAdd(retry_count, 1u)] = ti; return; }
    LfQView qv; LfTView tv;
    lf_task_views(d, task, qv, tv);

    uint32_t *smt = smem + tid;
    uint32_t qlo[NB], qhi[NB], qnn[NB], Pv[NB], Mv[NB];
#pragma unroll
    for (int w = 0; w < NB; w++) {
        if (w < nw) lf_q32(d, qv, (int64_t)w * 32, qlo[w], qhi[w], qnn[w]);
        else { qlo[w] = 0; qhi[w] = 0; qnn[w] = 0xffffffffu; }
        Pv[w] = 0xffffffffu; Mv[w] = 0u;
    }
    const int wl = (q - 1) >> 5;
    const uint32_t bl = (uint32_t)(q - 1) & 31u;
    int k = 0, top = 0;
    int cslide = lf_bandreg_next_slide<NB>(0, kmax, q, t);
    uint2 *ck = (uint2 *)(d.scratch + d.scr_off[ti]);

    /* ---- forward pass ---- */
    LfTStream ts;
    ts.init(d.pac, tv.t0, tv.dir);
    int c = 0;
    for (; c + C <= t; c += C) {
        if (c) {
            uint2 *dst = ck + (size_t)(c / C - 1) * NB;
#pragma unroll
            for (int w = 0; w < NB; w++) dst[w] = make_uint2(Pv[w], Mv[w]);
        }
        const uint32_t tb = ts.peek();
        ts.advance(C);
        lf_bandreg_block<NB, NB, true, false, true, SLIDE>(Pv, Mv, qlo, qhi, qnn, tb, c, C, k, cslide, top, q, t, kmax, d, qv, nullptr, 0);
    }
    if (c < t) {
        if (c) {
            uint2 *dst = ck + (size_t)(c / C - 1) * NB;
#pragma unroll
            for (int w = 0; w < NB; w++) dst[w] = make_uint2(Pv[w], Mv[w]);
        }
        lf_bandreg_block<NB, NB, true, false, false, SLIDE>(Pv, Mv, qlo, qhi, qnn, ts.peek(), c, t - c, k, cslide, top, q, t, kmax, d, qv, nullptr, 0);
    }
    int ed = t + top;
#pragma unroll
    for (int w = 0; w < NB; w++) {
        const int wa = k + w;
        const uint32_t m = wa < wl ? 0xffffffffu : wa == wl ? (0xffffffffu >> (31u - bl)) : 0u;
        ed += __popc(Pv[w] & m) - __popc(Mv[w] & m);
    }
    if (kmax > 0) {
        if (ed > cert) { LF_BAND_COUNT(lf_emu_band_retry); retry_list[first + atomicAdd(retry_count, 1u)] = ti; return; }
        LF_BAND_COUNT(lf_emu_band_ok);
    }
    lf_align_result r;
    r.edit_distance = ed; r.end_location = t - 1; r.status = 0;
    const uint64_t slot_hi = d.slot_end[ti] * 16ull;
    if (task.flags & LF_F_NO_PATH) { r.ops_off = slot_hi; r.ops_len = 0; d.res[ti] = r; return; }

    /* ---- traceback: Up > Left > Diagonal (edlib.cpp:950, :984, :1015) ---- */
    LfOpSink sink;
    sink.init(slot_hi);
    int i = q, j = t;
    const int q8 = (int)((8u * (uint32_t)q) / (uint32_t)t), r8 = (int)((8u * (uint32_t)q) % (uint32_t)t);
    int bc0 = ((t - 1) / C) * C;                 /* block the band bookkeeping below refers to */
    int rl0 = (int)(((uint32_t)bc0 * (uint32_t)q) / (uint32_t)t), ra0 = (int)(((uint32_t)bc0 * (uint32_t)q) % (uint32_t)t);
    int ks = -1, cs_of_ks = 0;                   /* memo of lf_bandreg_next_slide */
    bool lost = false;
    while (i > 0 && j > 0) {
        const int c1 = j, c0 = ((j - 1) / C) * C;
        while (bc0 > c0) { bc0 -= C; rl0 -= q8; ra0 -= r8; if (ra0 < 0) { ra0 += t; rl0--; } }
        int k0 = SLIDE ? (rl0 - 16 * NB + 16) >> 5 : 0;       /* band position after column c0-1: floor(c0*q/t) decides */
        k0 = k0 < 0 ? 0 : k0 > kmax ? kmax : k0;
        while (SLIDE && k != k0) { /* bring the query words of that band position into the registers (one word per ~32 columns) */
            if (k > k0) {
                k--;
#pragma unroll
                for (int w = NB - 1; w > 0; w--) { qlo[w] = qlo[w - 1]; qhi[w] = qhi[w - 1]; qnn[w] = qnn[w - 1]; }
                lf_q32(d, qv, (int64_t)k * 32, qlo[0], qhi[0], qnn[0]);
            } else {
#pragma unroll
                for (int w = 0; w + 1 < NB; w++) { qlo[w] = qlo[w + 1]; qhi[w] = qhi[w + 1]; qnn[w] = qnn[w + 1]; }
                k++;
                if (k + NB - 1 < nw) lf_q32(d, qv, (int64_t)(k + NB - 1) * 32, qlo[NB - 1], qhi[NB - 1], qnn[NB - 1]);
                else { qlo[NB - 1] = 0; qhi[NB - 1] = 0; qnn[NB - 1] = 0xffffffffu; }
            }
        }
        if (ks != k0) { ks = k0; cs_of_ks = lf_bandreg_next_slide<NB>(k0, kmax, q, t); }
        cslide = cs_of_ks;
        const int whi = (i - 1) >> 5;
        const int wtop = whi - WIN + 1 > 0 ? whi - WIN + 1 : 0;
        const int need = whi - k0 + 1;            /* band words down to the row the walk stands on */
        if (c0 == 0) {
#pragma unroll
            for (int w = 0; w < NB; w++) { Pv[w] = 0xffffffffu; Mv[w] = 0u; }
        } else {
            const uint2 *src = ck + (size_t)(c0 / C - 1) * NB;
#pragma unroll
            for (int w = 0; w < NB; w++) { if (w < need) { const uint2 v = src[w]; Pv[w] = v.x; Mv[w] = v.y; } }
        }
        ts.init(d.pac, tv.t0 + (int64_t)tv.dir * c0, tv.dir);
        const uint32_t tb = ts.peek();
        constexpr int N1 = NB > 2 ? NB - 1 : NB, N2 = NB > 3 ? NB - 2 : N1;
        const int ns = lf_converged_max(c1 - c0 != C ? NB : need);   /* one variant for the lanes that are here together */
        if (ns > N1) lf_bandreg_block<NB, NB, false, true, false, SLIDE>(Pv, Mv, qlo, qhi, qnn, tb, c0, c1 - c0, k, cslide, top, q, t, kmax, d, qv, smt, wtop);
        else if (N2 < N1 && ns <= N2) lf_bandreg_block<NB, N2, false, true, true, SLIDE>(Pv, Mv, qlo, qhi, qnn, tb, c0, C, k, cslide, top, q, t, kmax, d, qv, smt, wtop);
        else if (N1 < NB && ns <= N1) lf_bandreg_block<NB, N1, false, true, true, SLIDE>(Pv, Mv, qlo, qhi, qnn, tb, c0, C, k, cslide, top, q, t, kmax, d, qv, smt, wtop);
        else lf_bandreg_block<NB, NB, false, true, true, SLIDE>(Pv, Mv, qlo, qhi, qnn, tb, c0, C, k, cslide, top, q, t, kmax, d, qv, smt, wtop);
        /* walk inside the window, one word-row at a time */
        const int rowmin = wtop * 32;
        while (i > 0 && j > c0 && (i - 1) >= rowmin) {
            const int wrow = (i - 1) >> 5;
            lf_walk_row<128, CS>(smem, (int)tid + (wrow - wtop) * 2 * 128, wrow, c0, i, j, sink, d.ops);
        }
        if (sink.nops > (uint32_t)(q + t)) { lost = true; break; } /* cannot happen on a certified band */
    }
    while (i > 0 && !lost) { sink.emit(d.ops, 1u); i--; }
    while (j > 0 && !lost) { sink.emit(d.ops, 2u); j--; }
    sink.finish(d.ops);
    if (lost) { retry_list[first + atomicAdd(retry_count, 1u)] = ti; return; }
    const uint32_t nops = sink.nops;
    r.ops_off = slot_hi - nops; r.ops_len = nops;
    d.res[ti] = r;
}
