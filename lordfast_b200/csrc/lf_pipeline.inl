/*
 * lf_pipeline.inl -- host side of liblfgpu.so: the C ABI of include/lf_gpu.h on top of the kernels
 * in lf_kernels.cuh.  Included by lfgpu.cu (nvcc, the product) and by tests/emu/lfgpu_emu.cpp
 * (g++ + fiber emulator, test-only) so that the orchestration below is exercised in both.
 *
 * Per batch and per device:
 *   upload   reads (bytes + offsets) and tasks -> HBM; k_pack_reads builds the bit planes
 *   run      k_align_prep -> radix sort by (size class, target length) -> scans for op slots and
 *            checkpoint scratch -> one small D2H of the class histogram -> k_myers_small<NW,SHW>
 *            per non-empty class -> k_myers_large on a persistent grid of warp slots
 *   download results (24 B / task) and the 2-bit op stream
 * A batch is split over the context's devices by contiguous task ranges; the reference and the
 * reads are replicated, no collective is involved.
 */
#include <stddef.h>
#include <stdio.h>
#include <time.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "lf_gpu.h"
#include "lf_kernels.cuh"
#include "lf_backend.h"

#ifndef LF_NSUB
#define LF_NSUB 16   /* class streams per context.  24 was marginally better for one context alone (1.87 vs 1.88 ms per config-2 step), but two
                      * contexts in flight then hold 54 streams on the device's 32 hardware queues and their kernels falsely serialise:
                      * 14.6 ms per chunk with two lf_gpu_align_chains calls in flight, 10.7 ms with 16 streams each (profiles/r03) */
#endif
struct DevState {
    int dev = 0;
    lfb_stream stream = 0;
    lfb_stream ext_stream = 0;      /* speculative extensions of the chain operator run beside round 1 */
    lfb_event ext_ev = 0;
    lfb_event ev[4] = {};
    lfb_stream sub[LF_NSUB] = {};   /* size classes run concurrently: their grids are small */
    lfb_event sub_ev[LF_NSUB] = {};
    lfb_event cls_ev[LF_NCLS][2] = {};   /* start / end of every class kernel (timeline hook) */
    bool cls_ran[LF_NCLS] = {};
    uint32_t cls_count[LF_NCLS] = {};
    LfbBuf pac, bases, read_off, plo, phi, pnn;
    LfbBuf planes, gbytes, goff;   /* k_myers_band: plane regions per warp group */
    LfbBuf retry_scr;              /* warp slots of the k_myers_large instances that redo what k_myers_bandreg could not certify */
    LfbBuf group_scr;              /* plane scratch of the k_myers_group path classes, one region per class */
    LfbBuf res_keep, ops_keep;   /* lf_chain.inl parks the round-1 results / op stream here while round 3 runs */
    LfbBuf tasks, res, ops, keys, keys2, idx, idx2, slot_words, scr_bytes, slot_end, scr_off, scratch, large_scr, counters, queue;
    LfbBuf etasks, eres, escr_items, escr_off, escr;
    LfbTemp tmp;
    void *pinned = nullptr; /* LfCounters + two totals */
    uint32_t n_reads = 0; uint64_t total_bases = 0;
    size_t task_first = 0; uint32_t n_tasks = 0; uint64_t ops_words = 0;
    size_t etask_first = 0; uint32_t n_etasks = 0;
    bool ran = false;
    std::vector<std::pair<void *, size_t>> stage_chunks;   /* pinned staging for h2d_k from pageable sources, recycled per call */
    size_t stage_chunk = 0, stage_used = 0;
    bool pac_borrowed = false;      /* a lane of lf_gpu_align_chains: the reference belongs to the parent context's DevState */
    lfb_event up_ev = 0;            /* lane: its reads have arrived (recorded on the parent's upload stream) */
    lfb_event off_ev = 0;           /* the read offsets of the batch are on the device (the bases may still be in flight) */
    bool reads_preloaded = false;   /* lane: the parent has enqueued the H2D copy of the reads; upload_reads only waits and packs */
};

/* Kernel routing and scheduling switches: environment overrides read ONCE, when the context is created (a batch never
 * calls getenv); the defaults are the measured choices documented next to each reader below. */
struct LfRouting {
    uint32_t bandreg = 0, groupk = 0, band_mask = 0, group_wide = 0;
    bool bandreg_small = true, large_one_warp = false, serial = false, order_band_first = false;
    int nstreams = 0;
};
struct lf_gpu_ctx {
    LfRouting rt;
    std::vector<DevState> devs;
    std::vector<lf_gpu_ctx *> lanes;   /* lf_chain.inl: single-device child contexts that pipeline the sub-batches of one call */
    void *chain_scratch = nullptr; /* lf_chain.inl: pinned staging kept between calls */
    void (*chain_scratch_free)(void *) = nullptr;
    void *seed_state = nullptr;    /* lf_seed.inl: the FM index on device 0 and the buffers of lf_gpu_seed_batch */
    void (*seed_state_free)(void *) = nullptr;
    int64_t l_pac = 0;
    std::string err;
    lf_gpu_stats stats = {};
};

namespace {

struct HostTotals { LfCounters cnt; unsigned long long slot_total, scr_total; };

#ifndef LF_EMU
int set_dev(const DevState &d) { if (cudaSetDevice(d.dev) != cudaSuccess) return -2; lfb_cur_dev = d.dev; return 0; }
#else
int set_dev(const DevState &) { return 0; }
#endif

int fail(lf_gpu_ctx *c, int code, const char *msg)
{
    if (c) c->err = std::string(msg) + (lfb_errbuf[0] ? std::string(": ") + lfb_errbuf : std::string());
    return code;
}
#define LF_TRY(expr) do { int rc_ = (expr); if (rc_ != 0) return fail(ctx, rc_ == -5 ? LF_ERR_NOMEM : LF_ERR_CUDA, #expr); } while (0)

#ifndef LF_BANDREG_DEFAULT
#define LF_BANDREG_DEFAULT 0xf   /* NW = 6, 8, 12, 16 classes (round 2: with the uncertified retries no longer behind a slow warp kernel, the NW = 6 class gains too: 1.88 vs 1.92 ms per config-2 step) */
#endif
/* Size classes whose near-diagonal global tasks run in k_myers_bandreg (LF_BANDREG overrides; the other tasks of
 * 128 < q <= 512 run full width in k_myers_small, which is also the retry path of the tasks the band cannot certify) */
uint32_t bandreg_on()
{   /* bit i: size class 4+i (NW = 6, 8, 12, 16) */
    const char *e = getenv("LF_BANDREG");
    return e ? (uint32_t)strtoul(e, nullptr, 0) & 15u : (uint32_t)LF_BANDREG_DEFAULT;
}

/* k_myers_group classes in use (LF_GROUPK overrides): bit 0 distance-only tasks of 513 .. 8192 rows, bit 1 path tasks of
 * 513 .. 2048 rows (leaves), bit 2 path tasks of 257 .. 512 rows that are off the diagonal or in prefix mode.  0 sends
 * all of them where they went before: k_myers_large / k_myers_small<12|16>. */
uint32_t groupk_on()
{
    const char *e = getenv("LF_GROUPK");
    return e ? (uint32_t)strtoul(e, nullptr, 0) & 7u : 7u;
}

/* Global-mode tasks of q <= 128 in k_myers_bandreg (the band is the whole column: no slides, no certificate, nothing per
 * column in HBM) instead of k_myers_band (op planes of every column to HBM).  On since the recompute variant is chosen
 * per warp: 1.72 vs 1.77 ms (sv 0.0) and 2.57 vs 2.62 ms (sv 0.1) per config-2 step, and 3.9 GB less HBM traffic
 * (profiles/r02c section 6).  LF_BANDREG_SMALL=0 restores the plane store. */
bool bandreg_small()
{
    const char *e = getenv("LF_BANDREG_SMALL");
    return !e || atoi(e) != 0;
}

LfDev make_dev(lf_gpu_ctx *ctx, DevState &d)
{
    LfDev v;
    v.pac = d.pac.as<uint8_t>(); v.l_pac = ctx->l_pac;
    v.bases = d.bases.as<uint8_t>(); v.read_off = d.read_off.as<uint64_t>(); v.n_reads = d.n_reads;
    v.plo = d.plo.as<uint32_t>(); v.phi = d.phi.as<uint32_t>(); v.pnn = d.pnn.as<uint32_t>();
    v.tasks = d.tasks.as<lf_align_task>(); v.n_tasks = d.n_tasks;
    v.res = d.res.as<lf_align_result>();
    v.ops = d.ops.as<uint32_t>();
    v.slot_end = d.slot_end.as<uint64_t>();
    v.scr_off = d.scr_off.as<uint64_t>();
    v.scratch = d.scratch.as<uint8_t>();
    v.planes = d.planes.as<uint8_t>();
    v.bandreg = ctx->rt.bandreg;
    v.groupk = ctx->rt.groupk;
    return v;
}

size_t align_up(size_t v, size_t a);
LfLargeCfg large_cfg(size_t mq, size_t mt, size_t max_planes);

/* Pinned staging of one call: regions stay valid until stage_reset() at the start of the next call. */
void stage_reset(DevState &d) { d.stage_chunk = 0; d.stage_used = 0; }
void *stage_get(DevState &d, size_t n)
{
    n = (n + 63) & ~(size_t)63;
    while (d.stage_chunk < d.stage_chunks.size()) {
        auto &c = d.stage_chunks[d.stage_chunk];
        if (d.stage_used + n <= c.second) { void *p = (char *)c.first + d.stage_used; d.stage_used += n; return p; }
        d.stage_chunk++; d.stage_used = 0;
    }
    const size_t cap = n > ((size_t)4 << 20) ? n : ((size_t)4 << 20);
    void *p = lfb_host_alloc(cap);
    if (!p) return nullptr;
    d.stage_chunks.push_back({p, cap});
    d.stage_chunk = d.stage_chunks.size() - 1; d.stage_used = n;
    return p;
}
/* Upload of n bytes by kernel (k_copy16) on stream s.  src_pinned: src is pinned host memory, 16-byte aligned and readable
 * up to the next multiple of 16; else it is staged.  dst: a device buffer with 16 bytes of slack. */
int h2d_k(DevState &d, void *dst, const void *src, size_t n, lfb_stream s, bool src_pinned)
{
    if (!n) return 0;
    if (!src_pinned || ((uintptr_t)src & 15u)) {
        void *st = stage_get(d, n + 16);
        if (!st) return -5;
        memcpy(st, src, n);
        src = st;
    }
    const size_t n16 = (n + 15) >> 4;
    unsigned grid = (unsigned)((n16 + 255) / 256);
    if (grid > 296u) grid = 296u;
    LFB_LAUNCH(k_copy16, grid, 256, 0, s, (uint4 *)dst, (const uint4 *)src, n16);
    return 0;
}

template <int CI, bool SHW>
void launch_small(const LfDev &v, const uint32_t *order, uint32_t first, uint32_t count, lfb_stream s, const uint32_t *retry_count)
{
    constexpr int NW = CI == 0 ? 1 : CI == 1 ? 2 : CI == 2 ? 3 : CI == 3 ? 4 : CI == 4 ? 6 : CI == 5 ? 8 : CI == 6 ? 12 : 16;
    constexpr int WIN = NW < 2 ? 1 : 2;
    const size_t smem = (size_t)LF_K1_C * WIN * 2 * LF_K1_BLOCK * sizeof(uint32_t);
    const uint32_t grid = (count + LF_K1_BLOCK - 1) / LF_K1_BLOCK;
    auto kern = k_myers_small<NW, SHW>;
    LFB_LAUNCH(kern, grid, LF_K1_BLOCK, smem, s, v, order, first, count, retry_count);
}

template <int NB, bool SLIDE = true>
void launch_bandreg(const LfDev &v, const uint32_t *order, uint32_t first, uint32_t count, lfb_stream s, uint32_t *retry_list, uint32_t *retry_count, int nwmax)
{   /* shared memory: the window planes of 8 columns, as much as the other size-class kernels use */
    const size_t smem = (size_t)8 * (NB < 2 ? 1 : 2) * 2 * 128 * sizeof(uint32_t);
    auto kern = k_myers_bandreg<NB, SLIDE>;
    LFB_LAUNCH(kern, (count + 127) / 128, 128, smem, s, v, order, first, count, retry_list, retry_count, nwmax);
}

template <int NB, bool BANDED, bool SHW>
void launch_band(const LfDev &v, const uint32_t *order, uint32_t first, uint32_t count, uint32_t gbase, const unsigned long long *goff, lfb_stream s,
                 uint32_t *retry_list, uint32_t *retry_count)
{
    constexpr int WIN = NB < 2 ? 1 : 2;
    const size_t smem = (size_t)LF_BAND_C * WIN * 2 * 128 * sizeof(uint32_t);
    const uint32_t grid = (count + 127) / 128;
    auto kern = k_myers_band<NB, BANDED, SHW>;
    LFB_LAUNCH(kern, grid, 128, smem, s, v, order, first, count, gbase, goff, retry_list, retry_count);
}

template <int LANES, int WPL, bool PATH>
void launch_group(const LfDev &v, const uint32_t *order, uint32_t first, uint32_t count, lfb_stream s, const LfGroupRun &run, unsigned blocks)
{
    auto kern = k_myers_group<LANES, WPL, PATH>;
    LFB_LAUNCH(kern, blocks, 128, 0, s, v, order, first, count, run);
}
/* lanes per task, words per lane and path / distance-only of a k_myers_group class.  wide: a class with few tasks gives
 * every task a whole warp (or half of one) -- the step waits for its longest task, and 400 inversion-sized tasks on 8 lanes
 * each took 1.4 ms of a 1.9 ms step; with many tasks (the config-5 sweep) the narrow shape is the efficient one
 * (15 instead of 19-24 instructions per word-column). */
struct GroupShape { int lanes, wpl; bool path; };
GroupShape group_shape(int cls, bool wide)
{
    switch (cls) {
    case LF_CLS_GP16: return wide ? GroupShape{16, 1, true} : GroupShape{4, 4, true};
    case LF_CLS_GP32: return wide ? GroupShape{32, 1, true} : GroupShape{8, 4, true};
    case LF_CLS_GP64: return wide ? GroupShape{32, 2, true} : GroupShape{8, 8, true};
    case LF_CLS_GD32: return wide ? GroupShape{32, 1, false} : GroupShape{8, 4, false};
    case LF_CLS_GD64: return wide ? GroupShape{32, 2, false} : GroupShape{8, 8, false};
    case LF_CLS_GD128: return wide ? GroupShape{32, 4, false} : GroupShape{16, 8, false};
    default: return GroupShape{32, 8, false};
    }
}
void launch_group_class(int cls, bool wide, const LfDev &v, const uint32_t *order, uint32_t first, uint32_t count, lfb_stream s, const LfGroupRun &run, unsigned blocks)
{
    switch (cls) {
    case LF_CLS_GP16: if (wide) launch_group<16, 1, true>(v, order, first, count, s, run, blocks); else launch_group<4, 4, true>(v, order, first, count, s, run, blocks); break;
    case LF_CLS_GP32: if (wide) launch_group<32, 1, true>(v, order, first, count, s, run, blocks); else launch_group<8, 4, true>(v, order, first, count, s, run, blocks); break;
    case LF_CLS_GP64: if (wide) launch_group<32, 2, true>(v, order, first, count, s, run, blocks); else launch_group<8, 8, true>(v, order, first, count, s, run, blocks); break;
    case LF_CLS_GD32: if (wide) launch_group<32, 1, false>(v, order, first, count, s, run, blocks); else launch_group<8, 4, false>(v, order, first, count, s, run, blocks); break;
    case LF_CLS_GD64: if (wide) launch_group<32, 2, false>(v, order, first, count, s, run, blocks); else launch_group<8, 8, false>(v, order, first, count, s, run, blocks); break;
    case LF_CLS_GD128: if (wide) launch_group<32, 4, false>(v, order, first, count, s, run, blocks); else launch_group<16, 8, false>(v, order, first, count, s, run, blocks); break;
    case LF_CLS_GD256: launch_group<32, 8, false>(v, order, first, count, s, run, blocks); break;
    default: break;
    }
}

#ifndef LF_BAND_MASK_DEFAULT
#define LF_BAND_MASK_DEFAULT 0x0000
#endif
/* band width (32-row words) the class is run with by k_myers_band; 0 = full-width k_myers_small only */
int band_nb(int cls)
{
    const int sc = cls >> 1, shw = cls & 1;
    if (sc <= 3) return sc + 1;                 /* q <= 128: the whole column, no banding (NW and SHW) */
    if (shw) return 0;                          /* long prefix-mode tasks: full width */
    return sc == 4 ? 3 : sc == 7 ? 5 : 4;       /* q <= 192 / 256, 384 / 512 */
}

/* Which classes run k_myers_band (bit = class id); the others run the full-width k_myers_small.  Measured on B200
 * (profiles/r01g_band_mask_sweep.txt): keeping the op planes in HBM instead of recomputing them was a small win for
 * the unbanded classes (q <= 128) until k_myers_bandreg took their global-mode tasks; for the prefix-mode ones that
 * are left (1.5 % of the tasks) it makes no difference, and without any plane class the step saves a kernel, a scan
 * and a host round trip, so the default is none.  LF_BAND_MASK overrides (tests run 0xff / 0xffff). */
uint32_t band_mask()
{
    const char *e = getenv("LF_BAND_MASK");
    return e ? (uint32_t)strtoul(e, nullptr, 0) : (uint32_t)LF_BAND_MASK_DEFAULT;
}

LfRouting read_routing()
{
    LfRouting r;
    r.bandreg = bandreg_on(); r.groupk = groupk_on(); r.bandreg_small = bandreg_small(); r.band_mask = band_mask();
    r.large_one_warp = getenv("LF_LARGE_ONE_WARP") != nullptr;   /* one warp per large task instead of two */
    r.serial = getenv("LF_SERIAL") != nullptr;                    /* one class kernel at a time (per-class durations for profiling) */
    r.order_band_first = getenv("LF_ORDER_BAND_FIRST") != nullptr;
    r.nstreams = getenv("LF_STREAMS") ? atoi(getenv("LF_STREAMS")) : LF_NSUB - 1;   /* class kernels in flight at once */
    if (r.nstreams < 1 || r.nstreams > LF_NSUB - 1) r.nstreams = LF_NSUB - 1;
    /* class sizes up to this go wide; measured on the config-2 step: off (0) 1.87 ms, 2048 1.99 ms -- the wide shapes run 4x the
     * warps for the same columns and the step is bound by issue slots, not by the longest task */
    r.group_wide = getenv("LF_GROUP_WIDE") ? (uint32_t)atoi(getenv("LF_GROUP_WIDE")) : 0u;
    return r;
}

void launch_small_class(int cls, const LfRouting &rt, const LfDev &v, const uint32_t *order, uint32_t first, uint32_t count, uint32_t gbase,
                        const unsigned long long *goff, lfb_stream s, uint32_t *rl, uint32_t *rc, uint32_t bmask)
{   /* rl: retry list (indexed like `order`), rc: this class's retry counter */
    if (cls < 8 && !(cls & 1) && rt.bandreg_small) {   /* global-mode tasks of q <= 128: the band is the whole column (no slides, no certificate) */
        switch (cls) {
        case 0: launch_bandreg<1, false>(v, order, first, count, s, rl, rc, 1); return;
        case 2: launch_bandreg<2, false>(v, order, first, count, s, rl, rc, 2); return;
        case 4: launch_bandreg<3, false>(v, order, first, count, s, rl, rc, 3); return;
        case 6: launch_bandreg<4, false>(v, order, first, count, s, rl, rc, 4); return;
        }
    }
    if (cls < LF_CLS_LARGE && band_nb(cls) && (bmask >> cls & 1u)) {
        switch (cls) {
        case 0: launch_band<1, false, false>(v, order, first, count, gbase, goff, s, rl, rc); return;
        case 1: launch_band<1, false, true>(v, order, first, count, gbase, goff, s, rl, rc); return;
        case 2: launch_band<2, false, false>(v, order, first, count, gbase, goff, s, rl, rc); return;
        case 3: launch_band<2, false, true>(v, order, first, count, gbase, goff, s, rl, rc); return;
        case 4: launch_band<3, false, false>(v, order, first, count, gbase, goff, s, rl, rc); return;
        case 5: launch_band<3, false, true>(v, order, first, count, gbase, goff, s, rl, rc); return;
        case 6: launch_band<4, false, false>(v, order, first, count, gbase, goff, s, rl, rc); return;
        case 7: launch_band<4, false, true>(v, order, first, count, gbase, goff, s, rl, rc); return;
        /* banded first, then the full-width kernel over the dense list of tasks the band could not certify */
        case 8: launch_band<3, true, false>(v, order, first, count, gbase, goff, s, rl, rc); launch_small<4, false>(v, rl, first, count, s, rc); return;
        case 10: launch_band<4, true, false>(v, order, first, count, gbase, goff, s, rl, rc); launch_small<5, false>(v, rl, first, count, s, rc); return;
        case 12: launch_band<4, true, false>(v, order, first, count, gbase, goff, s, rl, rc); launch_small<6, false>(v, rl, first, count, s, rc); return;
        case 14: launch_band<5, true, false>(v, order, first, count, gbase, goff, s, rl, rc); launch_small<7, false>(v, rl, first, count, s, rc); return;
        default: break;
        }
    }
    switch (cls) {
    /* sliding band in registers; the caller then runs k_myers_large over the dense list of tasks the band could not certify */
    case LF_CLS_BANDREG0 + 0: launch_bandreg<3>(v, order, first, count, s, rl, rc, 6); break;
    case LF_CLS_BANDREG0 + 1: launch_bandreg<4>(v, order, first, count, s, rl, rc, 8); break;
    case LF_CLS_BANDREG0 + 2: launch_bandreg<4>(v, order, first, count, s, rl, rc, 12); break;
    case LF_CLS_BANDREG0 + 3: launch_bandreg<5>(v, order, first, count, s, rl, rc, 16); break;
    case 0: launch_small<0, false>(v, order, first, count, s, nullptr); break;
    case 1: launch_small<0, true>(v, order, first, count, s, nullptr); break;
    case 2: launch_small<1, false>(v, order, first, count, s, nullptr); break;
    case 3: launch_small<1, true>(v, order, first, count, s, nullptr); break;
    case 4: launch_small<2, false>(v, order, first, count, s, nullptr); break;
    case 5: launch_small<2, true>(v, order, first, count, s, nullptr); break;
    case 6: launch_small<3, false>(v, order, first, count, s, nullptr); break;
    case 7: launch_small<3, true>(v, order, first, count, s, nullptr); break;
    case 8: launch_small<4, false>(v, order, first, count, s, nullptr); break;
    case 9: launch_small<4, true>(v, order, first, count, s, nullptr); break;
    case 10: launch_small<5, false>(v, order, first, count, s, nullptr); break;
    case 11: launch_small<5, true>(v, order, first, count, s, nullptr); break;
    case 12: launch_small<6, false>(v, order, first, count, s, nullptr); break;
    case 13: launch_small<6, true>(v, order, first, count, s, nullptr); break;
    case 14: launch_small<7, false>(v, order, first, count, s, nullptr); break;
    case 15: launch_small<7, true>(v, order, first, count, s, nullptr); break;
    default: break;
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

LfLargeCfg large_cfg(size_t mq, size_t mt, size_t max_planes)
{   /* layout of one warp slot of k_myers_large for tasks of up to mq x mt */
    LfLargeCfg cfg;
    size_t off = align_up(max_planes + 256, 256);
    cfg.off_hb = off; off = align_up(off + mt + 64, 256);
    cfg.off_L = off; off = align_up(off + (mq + 2) * 4, 256);
    cfg.off_R = off; off = align_up(off + (mq + 2) * 4, 256);
    cfg.off_opsb = off; off = align_up(off + mq + mt + 64, 256);
    cfg.off_stack = off; off = align_up(off + LF_LARGE_STACK * 5 * 4, 256);
    cfg.stride = off;
    cfg.base = nullptr; cfg.queue = nullptr;
    return cfg;
}

int run_align_dev(lf_gpu_ctx *ctx, DevState &d)
{
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    d.ran = false;
    d.ops_words = 0;
    if (d.n_tasks == 0) { d.ran = true; return LF_OK; }
    const size_t n = d.n_tasks;
    lfb_stream s = d.stream;
    LF_TRY(d.keys.reserve(n * 4)); LF_TRY(d.keys2.reserve(n * 4)); LF_TRY(d.idx.reserve(n * 4)); LF_TRY(d.idx2.reserve(n * 4));
    LF_TRY(d.slot_words.reserve(n * 4)); LF_TRY(d.scr_bytes.reserve(n * 4));
    LF_TRY(d.slot_end.reserve(n * 8)); LF_TRY(d.scr_off.reserve((n + 1) * 8));
    LF_TRY(d.res.reserve(n * sizeof(lf_align_result)));
    LF_TRY(d.counters.reserve(sizeof(LfCounters))); LF_TRY(d.queue.reserve(256)); /* [0]: large-task work counter, [1+cls]: retry counters */
    LF_TRY(lfb_memset(d.counters.p, 0, sizeof(LfCounters), s));
    LF_TRY(lfb_memset(d.queue.p, 0, 256, s));
#ifndef LF_EMU
    cudaEventRecord(d.ev[0], s);
#endif
    LfDev v = make_dev(ctx, d);
    LFB_LAUNCH(k_align_prep, (unsigned)((n + 255) / 256), 256, 0, s, v, d.keys.as<uint32_t>(), d.idx.as<uint32_t>(),
               d.slot_words.as<uint32_t>(), d.scr_bytes.as<uint32_t>(), d.counters.as<LfCounters>());
    LF_TRY(lfb_sort_pairs(d.tmp, d.keys.as<uint32_t>(), d.keys2.as<uint32_t>(), d.idx.as<uint32_t>(), d.idx2.as<uint32_t>(), n, s));
    LF_TRY(lfb_scan_incl(d.tmp, d.slot_words.as<uint32_t>(), d.slot_end.as<unsigned long long>(), n, s));
    LF_TRY(lfb_scan_excl_total(d.tmp, d.scr_bytes.as<uint32_t>(), d.scr_off.as<unsigned long long>(), n, s));
    HostTotals *ht = (HostTotals *)d.pinned;
    static_assert(sizeof(LfCounters) % 8 == 0 && offsetof(HostTotals, cnt) == 0, "k_readback writes HostTotals in place");
    LFB_LAUNCH(k_readback, 1, 64, 0, s, (const uint32_t *)d.counters.p, (uint32_t)(sizeof(LfCounters) / 4), d.slot_end.as<unsigned long long>() + (n - 1), d.scr_off.as<unsigned long long>() + n,
               (uint32_t *)ht, (uint32_t)(offsetof(HostTotals, slot_total) / 8), (uint32_t)(offsetof(HostTotals, scr_total) / 8));
    LF_TRY(lfb_sync(s)); /* the one host round trip of a batch: class sizes decide the launches */
    if (ht->slot_total >> 32) return fail(ctx, LF_ERR_BAD_ARG, "batch needs 2^32 or more op words: split it");   /* the kernels index op words with 32 bits */
    d.ops_words = ht->slot_total;
    LF_TRY(d.ops.reserve((size_t)ht->slot_total * 4 + 64));
    LF_TRY(d.scratch.reserve((size_t)ht->scr_total + 64));
    /* plane regions of k_myers_band: one per warp group (32 consecutive sorted tasks of a class) */
    LfGroupCfg gc;
    const LfRouting &rt = ctx->rt;
    const uint32_t bmask = rt.band_mask;
    {
        uint32_t first = 0, g = 0;
        for (int cls = 0; cls < LF_CLS_LARGE; cls++) {
            gc.first[cls] = first; gc.count[cls] = ht->cnt.hist[cls]; gc.gbase[cls] = g; gc.nb[cls] = ((bmask >> cls & 1u) && !(rt.bandreg_small && cls < 8 && !(cls & 1))) ? (uint32_t)band_nb(cls) : 0u;
            first += ht->cnt.hist[cls]; g += (ht->cnt.hist[cls] + 31) / 32;
        }
        gc.gbase[LF_CLS_LARGE] = g;
    }
    uint32_t ngroups = gc.gbase[LF_CLS_LARGE];
    { uint32_t any = 0; for (int cls = 0; cls < LF_CLS_LARGE; cls++) any |= gc.count[cls] ? gc.nb[cls] : 0u; if (!any) ngroups = 0; }   /* no class stores op planes: no regions, no second host round trip */
    if (ngroups) {
        LF_TRY(d.gbytes.reserve((size_t)ngroups * 4 + 64)); LF_TRY(d.goff.reserve(((size_t)ngroups + 1) * 8));
        v = make_dev(ctx, d);
        LFB_LAUNCH(k_group_scratch, (ngroups + 255) / 256, 256, 0, s, v, d.idx2.as<uint32_t>(), gc, d.gbytes.as<uint32_t>());
        LF_TRY(lfb_scan_excl_total(d.tmp, d.gbytes.as<uint32_t>(), d.goff.as<unsigned long long>(), ngroups, s));
        LFB_LAUNCH(k_readback, 1, 64, 0, s, (const uint32_t *)nullptr, 0u, (const unsigned long long *)nullptr, d.goff.as<unsigned long long>() + ngroups,
                   (uint32_t *)ht, 0u, (uint32_t)(offsetof(HostTotals, scr_total) / 8));
        LF_TRY(lfb_sync(s));
        LF_TRY(d.planes.reserve((size_t)ht->scr_total + 256));
    }
    v = make_dev(ctx, d);

    ctx->stats.align_tasks += n;
    ctx->stats.cells += ht->cnt.cells;
    ctx->stats.word_columns += ht->cnt.word_columns;
    ctx->stats.last_main_word_columns = ht->cnt.word_columns; /* all alignment kernels of the step run concurrently */

#ifndef LF_EMU
    cudaEventRecord(d.ev[2], s);
    for (int k = 0; k < LF_NSUB; k++) cudaStreamWaitEvent(d.sub[k], d.ev[2], 0);
#endif
    /* the large-task kernel first, on its own stream: few warps, long latency, overlaps everything else */
    uint32_t nsmall = 0;
    for (int cls = 0; cls < LF_CLS_LARGE; cls++) nsmall += ht->cnt.hist[cls];
    const uint32_t nlarge = ht->cnt.hist[LF_CLS_LARGE];
    if (nlarge) {
        LfLargeCfg cfg = large_cfg(ht->cnt.max_q, ht->cnt.max_t, ht->cnt.max_planes);
        /* persistent grid of warp slots; bounded so that the scratch stays within a few GB */
        /* blocks of two warps (two scratch slots each): a task above edlib's size rule is split once by both */
        size_t slots = 148 * 8;
        const size_t budget = (size_t)8 << 30;
        if (slots * 2 * cfg.stride > budget) slots = budget / (2 * cfg.stride);
        if (slots < 1) slots = 1;
        if (slots > nlarge) slots = nlarge;
        LF_TRY(d.large_scr.reserve(slots * 2 * cfg.stride));
        cfg.base = d.large_scr.as<uint8_t>();
        cfg.queue = d.queue.as<uint32_t>();
#ifndef LF_EMU
        cudaEventRecord(d.cls_ev[LF_CLS_LARGE][0], d.sub[0]);
#endif
        LFB_LAUNCH(k_myers_large, (unsigned)slots, rt.large_one_warp ? 32 : 64, 0, d.sub[0], v, d.idx2.as<uint32_t>(), nsmall, nlarge, cfg, (const uint32_t *)nullptr);
#ifndef LF_EMU
        cudaEventRecord(d.cls_ev[LF_CLS_LARGE][1], d.sub[0]);
#endif
    }
    for (int c = 0; c < LF_NCLS; c++) d.cls_count[c] = ht->cnt.hist[c];
    for (int c = 0; c < LF_NCLS; c++) d.cls_ran[c] = c == LF_CLS_BAD ? false : c == LF_CLS_LARGE ? nlarge != 0 : ht->cnt.hist[c] != 0;
    /* thread-per-task classes, the ones with the longest tasks first, spread over the other streams */
    {
        uint32_t firsts[LF_NCLS];
        uint32_t first = 0;
        for (int cls = 0; cls < LF_NCLS; cls++) { firsts[cls] = first; first += ht->cnt.hist[cls]; }   /* sorted order = class id order */
        int k = 0;
        int seq[LF_NCLS], nseq = 0;
        const bool serial = rt.serial;
        const int nstreams = rt.nstreams;
        /* k_myers_group classes first (their tasks are the longest after the large ones), longest rows first */
        for (int cls = LF_CLS_GD256; cls >= LF_CLS_GP16; cls--) seq[nseq++] = cls;
        if (!rt.order_band_first) {   /* the full-width classes (few, long tasks: the tail of the step) before the sliding-band ones */
            for (int cls = LF_CLS_LARGE - 1; cls >= 8; cls--) seq[nseq++] = cls;
            for (int cls = LF_CLS_BANDREG0 + 3; cls >= LF_CLS_BANDREG0; cls--) seq[nseq++] = cls;
            for (int cls = 7; cls >= 0; cls--) seq[nseq++] = cls;
        } else {
            for (int cls = LF_CLS_BANDREG0 + 3; cls >= LF_CLS_BANDREG0; cls--) seq[nseq++] = cls;
            for (int cls = LF_CLS_LARGE - 1; cls >= 0; cls--) seq[nseq++] = cls;
        }
        /* plane scratch of the group path classes: one slot per group of lanes of the persistent grid */
        unsigned gblocks[LF_NCLS] = {};
        bool gwide[LF_NCLS] = {};
        LfGroupRun grun[LF_NCLS] = {};
        const uint32_t wide_below = rt.group_wide;
        {
            size_t off = 0, offs[LF_NCLS] = {};
            for (int cls = LF_CLS_GP16; cls <= LF_CLS_GD256; cls++) {
                const uint32_t count = ht->cnt.hist[cls];
                if (!count) continue;
                gwide[cls] = count <= wide_below;
                const GroupShape gs = group_shape(cls, gwide[cls]);
                const unsigned per_block = 4u * (32u / (unsigned)gs.lanes);   /* tasks a block works on at a time */
                unsigned blocks = (count + per_block - 1) / per_block;
                if (blocks > 148u * 6u) blocks = 148u * 6u;
                size_t stride = 0;
                if (gs.path) {
                    stride = align_up((size_t)ht->cnt.gmax_t[cls - LF_CLS_GP16] * (size_t)(gs.lanes * gs.wpl) * 8 + 256, 256);
                    const size_t budget = (size_t)3 << 30;
                    while (blocks > 1 && (size_t)blocks * per_block * stride > budget) blocks = (blocks + 1) / 2;
                }
                gblocks[cls] = blocks; grun[cls].stride = stride; offs[cls] = off;
                off += (size_t)blocks * per_block * stride;
            }
            if (off) LF_TRY(d.group_scr.reserve(off + 256));
            for (int cls = LF_CLS_GP16; cls <= LF_CLS_GD256; cls++) {
                grun[cls].base = d.group_scr.as<uint8_t>() + offs[cls];
                grun[cls].queue = d.queue.as<uint32_t>() + 40 + (cls - LF_CLS_GP16);
            }
        }
        for (int si = 0; si < nseq; si++) {
            const int cls = seq[si];
            const uint32_t count = ht->cnt.hist[cls];
            if (!count) continue;
            lfb_stream st = serial ? d.sub[1] : d.sub[1 + (k++ % nstreams)];   /* LF_SERIAL=1: one class at a time (per-class durations for profiling) */
#ifndef LF_EMU
            cudaEventRecord(d.cls_ev[cls][0], st);
#endif
            if (cls >= LF_CLS_GP16) launch_group_class(cls, gwide[cls], v, d.idx2.as<uint32_t>(), firsts[cls], count, st, grun[cls], gblocks[cls]);
            else launch_small_class(cls, rt, v, d.idx2.as<uint32_t>(), firsts[cls], count, cls < LF_CLS_LARGE ? gc.gbase[cls] : 0u, d.goff.as<unsigned long long>(), st,
                               d.idx.as<uint32_t>() /* input of the sort, free by now */, d.queue.as<uint32_t>() + 1 + cls, bmask);
            if (cls >= LF_CLS_BANDREG0 && cls < LF_CLS_GP16) {   /* the uncertified few: warp per task, on the same stream */
                const int bi = cls - LF_CLS_BANDREG0;
                LfLargeCfg rcfg = large_cfg(512, 640, (size_t)lf_large_planes_bytes(512, 640));   /* eligible tasks: q <= 512, |q-t| <= 40 */
                const size_t rslots = 148;
                LF_TRY(d.retry_scr.reserve(4 * rslots * rcfg.stride));
                rcfg.base = d.retry_scr.as<uint8_t>() + (size_t)bi * rslots * rcfg.stride;
                rcfg.queue = d.queue.as<uint32_t>() + 32 + bi;
                LFB_LAUNCH(k_myers_large, (unsigned)rslots, 32, 0, st, v, d.idx.as<uint32_t>(), firsts[cls], count, rcfg, (const uint32_t *)(d.queue.as<uint32_t>() + 1 + cls));
            }
#ifndef LF_EMU
            cudaEventRecord(d.cls_ev[cls][1], st);
#endif
        }
    }
#ifndef LF_EMU
    for (int k = 0; k < LF_NSUB; k++) { cudaEventRecord(d.sub_ev[k], d.sub[k]); cudaStreamWaitEvent(s, d.sub_ev[k], 0); }
    cudaEventRecord(d.ev[3], s);
    cudaEventRecord(d.ev[1], s);
#endif
    LF_TRY(lfb_last_error());
    d.ran = true;
    return LF_OK;
}

} // namespace

/* ---------------------------------------------------------------------------------------------- */
#ifndef LF_EMU
static std::mutex g_prewarm_mu;
static std::thread g_prewarm;
static bool g_prewarm_started = false;
/* a process that calls lf_gpu_prewarm() and exits without ever reaching lf_gpu_init() (bad command line, empty input) must
 * not die in std::terminate() on the still-joinable thread: join it when the library's statics are torn down */
static struct PrewarmGuard { ~PrewarmGuard() { std::lock_guard<std::mutex> g(g_prewarm_mu); if (g_prewarm.joinable()) g_prewarm.join(); } } g_prewarm_guard;
#endif

/* streams and events of one device state (the caller has made d.dev current) */
static int init_dev_streams(DevState &d)
{
#ifndef LF_EMU
    /* Three priority levels.  Highest: the extension stream (its few long-latency warps must not queue behind the round-1
     * grids).  Middle: the main stream and the class streams -- above the default level, at which the chain operator's
     * early emit runs: the blocks of a kernel launched later are only dispatched once every block of the earlier kernels of
     * the same level has been, so round 3's first tiny kernel used to wait ~2 ms for the early emit's 18 000 blocks. */
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const int mid = (hi < lo && !(getenv("LF_STREAM_PRIO") && atoi(getenv("LF_STREAM_PRIO")) == 0)) ? lo - 1 : lo;   /* numerically lower = more urgent; LF_STREAM_PRIO=0: one level for everything but the extensions */
    if (cudaStreamCreateWithPriority(&d.stream, cudaStreamNonBlocking, mid) != cudaSuccess) return -1;
    for (int k = 0; k < 4; k++) cudaEventCreate(&d.ev[k]);
    if (cudaStreamCreateWithPriority(&d.ext_stream, cudaStreamNonBlocking, hi) != cudaSuccess) return -1;
    cudaEventCreateWithFlags(&d.ext_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&d.up_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&d.off_ev, cudaEventDisableTiming);
    for (int k = 0; k < LF_NSUB; k++) { if (cudaStreamCreateWithPriority(&d.sub[k], cudaStreamNonBlocking, mid) != cudaSuccess) return -1; cudaEventCreateWithFlags(&d.sub_ev[k], cudaEventDisableTiming); }
    for (int c = 0; c < LF_NCLS; c++) { cudaEventCreate(&d.cls_ev[c][0]); cudaEventCreate(&d.cls_ev[c][1]); }
#else
    (void)d;
#endif
    return 0;
}

extern "C" {

int lf_gpu_init(lf_gpu_ctx **out, const uint8_t *pac, int64_t l_pac, const int *devices, int n_devices)
{
    if (!out || !pac || l_pac <= 0) return LF_ERR_BAD_ARG;
#ifndef LF_EMU
    const bool init_trace = getenv("LF_INIT_TRACE") != nullptr;
    struct timespec its0; clock_gettime(CLOCK_MONOTONIC, &its0);
    { std::lock_guard<std::mutex> g(g_prewarm_mu); if (g_prewarm.joinable()) g_prewarm.join(); }
    if (init_trace) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); fprintf(stderr, "[lf_gpu_init: waited %.0f ms for the prewarm thread]\n", (ts.tv_sec - its0.tv_sec) * 1e3 + (ts.tv_nsec - its0.tv_nsec) * 1e-6); }
#endif
    *out = nullptr;
    lfb_errbuf[0] = 0;
    std::vector<int> devs;
#ifndef LF_EMU
    /* one hardware queue per class stream (the default of 8 connections serialises 18 streams in pairs);
     * only effective if no CUDA context exists yet -- bench.py / api.py also set it before CUDA starts */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return LF_ERR_NO_DEVICE;
    if (!devices || n_devices <= 0) { int cur = 0; if (cudaGetDevice(&cur) != cudaSuccess) return LF_ERR_NO_DEVICE; devs.push_back(cur); }
    else for (int i = 0; i < n_devices; i++) { if (devices[i] < 0 || devices[i] >= count) return LF_ERR_BAD_ARG; devs.push_back(devices[i]); }
#else
    (void)devices; (void)n_devices;
    devs.push_back(0);
#endif
    lf_gpu_ctx *ctx = new lf_gpu_ctx();
    ctx->rt = read_routing();
    ctx->l_pac = l_pac;
    ctx->devs.resize(devs.size());
    const size_t pac_bytes = (size_t)(l_pac / 4 + 1);
    for (size_t i = 0; i < devs.size(); i++) {
        DevState &d = ctx->devs[i];
        d.dev = devs[i];
        if (set_dev(d)) { delete ctx; return LF_ERR_NO_DEVICE; }
        if (init_dev_streams(d)) { delete ctx; return LF_ERR_CUDA; }
        d.pinned = lfb_host_alloc(sizeof(HostTotals));
        if (!d.pinned || d.pac.reserve(pac_bytes + 16)) { lf_gpu_destroy(ctx); return LF_ERR_NOMEM; }
        if (lfb_memset(d.pac.p, 0, pac_bytes + 16, d.stream) || lfb_h2d(d.pac.p, pac, pac_bytes, d.stream) || lfb_sync(d.stream)) { lf_gpu_destroy(ctx); return LF_ERR_CUDA; }
    }
#ifndef LF_EMU
    if (init_trace) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); fprintf(stderr, "[lf_gpu_init: %.0f ms in all (streams, events, reference upload after the wait)]\n", (ts.tv_sec - its0.tv_sec) * 1e3 + (ts.tv_nsec - its0.tv_nsec) * 1e-6); }
#endif
    *out = ctx;
    return LF_OK;
}

void lf_gpu_destroy(lf_gpu_ctx *ctx)
{
    if (!ctx) return;
    for (lf_gpu_ctx *l : ctx->lanes) lf_gpu_destroy(l);
    ctx->lanes.clear();
    if (ctx->chain_scratch && ctx->chain_scratch_free) ctx->chain_scratch_free(ctx->chain_scratch);
    if (ctx->seed_state && ctx->seed_state_free) { if (!ctx->devs.empty()) set_dev(ctx->devs[0]); ctx->seed_state_free(ctx->seed_state); }
    for (DevState &d : ctx->devs) {
        set_dev(d);
        if (d.pac_borrowed) { d.pac.p = nullptr; d.pac.cap = 0; }
        LfbBuf *bufs[] = { &d.retry_scr, &d.group_scr, &d.planes, &d.gbytes, &d.goff, &d.res_keep, &d.ops_keep, &d.pac, &d.bases, &d.read_off, &d.plo, &d.phi, &d.pnn, &d.tasks, &d.res, &d.ops, &d.keys, &d.keys2, &d.idx, &d.idx2,
                           &d.slot_words, &d.scr_bytes, &d.slot_end, &d.scr_off, &d.scratch, &d.large_scr, &d.counters, &d.queue,
                           &d.etasks, &d.eres, &d.escr_items, &d.escr_off, &d.escr };
        for (LfbBuf *b : bufs) b->release();
        lfb_free(d.tmp.p);
        lfb_host_free(d.pinned);
        for (auto &c : d.stage_chunks) lfb_host_free(c.first);
        d.stage_chunks.clear();
#ifndef LF_EMU
        for (int k = 0; k < 4; k++) if (d.ev[k]) cudaEventDestroy(d.ev[k]);
        for (int k = 0; k < LF_NSUB; k++) { if (d.sub_ev[k]) cudaEventDestroy(d.sub_ev[k]); if (d.sub[k]) cudaStreamDestroy(d.sub[k]); }
        if (d.ext_ev) cudaEventDestroy(d.ext_ev);
        if (d.up_ev) cudaEventDestroy(d.up_ev);
        if (d.off_ev) cudaEventDestroy(d.off_ev);
        if (d.ext_stream) cudaStreamDestroy(d.ext_stream);
        for (int c = 0; c < LF_NCLS; c++) { if (d.cls_ev[c][0]) cudaEventDestroy(d.cls_ev[c][0]); if (d.cls_ev[c][1]) cudaEventDestroy(d.cls_ev[c][1]); }
        if (d.stream) cudaStreamDestroy(d.stream);
#endif
    }
    delete ctx;
}

const char *lf_gpu_last_error(const lf_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void lf_gpu_prewarm(void)
{
#ifndef LF_EMU
    std::lock_guard<std::mutex> g(g_prewarm_mu);
    if (g_prewarm_started) return;
    g_prewarm_started = true;
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    g_prewarm = std::thread([] {
        const bool tr = getenv("LF_INIT_TRACE") != nullptr;
        auto now = [] { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
        const double t0 = now();
        cudaFree(nullptr);                                  /* driver + primary context */
        const double t1 = now();
        cudaFuncAttributes a;
        cudaFuncGetAttributes(&a, (const void *)k_pack_reads);   /* module load */
        const double t2 = now();
        void *p = nullptr;                                  /* first pinned allocation maps the host-memory machinery */
        if (cudaMallocHost(&p, 1 << 20) == cudaSuccess) cudaFreeHost(p);
        if (tr) fprintf(stderr, "[lf_gpu_prewarm: driver + context %.0f ms, module %.0f ms, first pinned allocation %.0f ms]\n", t1 - t0, t2 - t1, now() - t2);
    });
#endif
}

void *lf_gpu_host_alloc(size_t bytes) { return lfb_host_alloc(bytes); }
void lf_gpu_host_free(void *p) { lfb_host_free(p); }

size_t lf_gpu_ops_capacity(const lf_align_task *tasks, size_t n)
{
    size_t words = 0;
    for (size_t i = 0; i < n; i++)
        if (!(tasks[i].flags & LF_F_NO_PATH)) words += ((size_t)tasks[i].q_len + tasks[i].t_len + 15) >> 4;
    return words * 4 + 64;
}

int lf_gpu_upload_reads(lf_gpu_ctx *ctx, const lf_reads *reads)
{
    if (!ctx || !reads || !reads->offsets || (!reads->bases && reads->n_reads)) return LF_ERR_BAD_ARG;
    const uint32_t nr = reads->n_reads;
    const uint64_t total = reads->offsets[nr] - reads->offsets[0];
    const size_t pwords = (size_t)(total >> 5) + 3 * (size_t)nr + 8;
    for (DevState &d : ctx->devs) {
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        stage_reset(d);   /* a new batch starts here */
        LF_TRY(d.plo.reserve(pwords * 4)); LF_TRY(d.phi.reserve(pwords * 4)); LF_TRY(d.pnn.reserve(pwords * 4));
        if (d.reads_preloaded) {   /* lf_chain.inl lanes: the parent enqueued the copies in lane order on its upload stream */
#ifndef LF_EMU
            cudaStreamWaitEvent(d.stream, d.up_ev, 0);
            cudaEventRecord(d.off_ev, d.stream);
#endif
            d.reads_preloaded = false;
        } else {
            if (reads->offsets[0] != 0) return fail(ctx, LF_ERR_BAD_ARG, "read offsets must start at 0");
            LF_TRY(d.bases.reserve(total + 64)); LF_TRY(d.read_off.reserve(((size_t)nr + 1) * 8 + 64));
            /* the offsets first and through pinned staging: a cudaMemcpyAsync from pageable memory blocks the calling thread
             * (and, it turned out, other threads' CUDA calls) until everything queued on the stream before it -- 200 MB of
             * bases -- has gone through */
            LF_TRY(h2d_k(d, d.read_off.p, reads->offsets, ((size_t)nr + 1) * 8, d.stream, false));
#ifndef LF_EMU
            cudaEventRecord(d.off_ev, d.stream);
#endif
            LF_TRY(lfb_h2d(d.bases.p, reads->bases, total, d.stream));
        }
        LF_TRY(lfb_memset(d.plo.p, 0, pwords * 4, d.stream)); LF_TRY(lfb_memset(d.phi.p, 0, pwords * 4, d.stream));
        LF_TRY(lfb_memset(d.pnn.p, 0xff, pwords * 4, d.stream));
        d.n_reads = nr; d.total_bases = total;
        if (nr) {
            LfDev v = make_dev(ctx, d);
            LFB_LAUNCH(k_pack_reads, nr, 128, 0, d.stream, v);
        }
    }
    return LF_OK;
}

int lf_gpu_pack_reads(lf_gpu_ctx *ctx)
{
    if (!ctx) return LF_ERR_BAD_ARG;
    for (DevState &d : ctx->devs) {
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        if (!d.n_reads) continue;
        LfDev v = make_dev(ctx, d);
        LFB_LAUNCH(k_pack_reads, d.n_reads, 128, 0, d.stream, v);
    }
    return LF_OK;
}

int lf_gpu_upload_align_tasks(lf_gpu_ctx *ctx, const lf_align_task *tasks, size_t n)
{
    if (!ctx || (!tasks && n) || n > 0x7fffffffu) return LF_ERR_BAD_ARG;
    const size_t nd = ctx->devs.size();
    for (size_t k = 0; k < nd; k++) {
        DevState &d = ctx->devs[k];
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        const size_t lo = n * k / nd, hi = n * (k + 1) / nd;
        d.task_first = lo; d.n_tasks = (uint32_t)(hi - lo); d.ran = false;
        LF_TRY(d.tasks.reserve((hi - lo) * sizeof(lf_align_task)));
        if (hi > lo) LF_TRY(lfb_h2d(d.tasks.p, tasks + lo, (hi - lo) * sizeof(lf_align_task), d.stream));
    }
    return LF_OK;
}

/* Chain operator, single-device contexts: the round-1 tasks are written by k_chain_tasks straight into the device's
 * task array; this sizes it and makes it the resident batch. */
static lf_align_task *resident_tasks_alloc(lf_gpu_ctx *ctx, size_t n)
{
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return nullptr;
    if (d.tasks.reserve((n + 1) * sizeof(lf_align_task))) return nullptr;
    d.task_first = 0; d.n_tasks = (uint32_t)n; d.ran = false;
    return d.tasks.as<lf_align_task>();
}

/* Chain operator, single-device contexts: the follow-up (round-3) tasks, uploaded by kernel (see k_copy16). */
static int upload_align_tasks_k(lf_gpu_ctx *ctx, const lf_align_task *tasks, size_t n)
{
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    d.task_first = 0; d.n_tasks = (uint32_t)n; d.ran = false;
    LF_TRY(d.tasks.reserve(n * sizeof(lf_align_task) + 64));
    LF_TRY(h2d_k(d, d.tasks.p, tasks, n * sizeof(lf_align_task), d.stream, false));
    return LF_OK;
}

int lf_gpu_run_align(lf_gpu_ctx *ctx)
{
    if (!ctx) return LF_ERR_BAD_ARG;
    /* the per-batch host sync inside run_align_dev serialises devices of one context; callers that
     * want several GPUs busy use one context (one process) per GPU, as bench.py does */
    for (DevState &d : ctx->devs) { int rc = run_align_dev(ctx, d); if (rc) return rc; }
    ctx->stats.kernel_launches = lfb_launches;
    return LF_OK;
}

int lf_gpu_sync(lf_gpu_ctx *ctx)
{
    if (!ctx) return LF_ERR_BAD_ARG;
    for (DevState &d : ctx->devs) { if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice"); LF_TRY(lfb_sync(d.stream)); }
#ifndef LF_EMU
    DevState &d0 = ctx->devs[0];
    if (d0.ran && d0.n_tasks) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, d0.ev[0], d0.ev[1]) == cudaSuccess) ctx->stats.last_run_ms = ms;
        if (cudaEventElapsedTime(&ms, d0.ev[2], d0.ev[3]) == cudaSuccess) ctx->stats.last_main_kernel_ms = ms;
    }
#endif
    return LF_OK;
}

int lf_gpu_download_align(lf_gpu_ctx *ctx, lf_align_result *res, uint8_t *ops, size_t ops_cap)
{
    if (!ctx || !res) return LF_ERR_BAD_ARG;
    uint64_t base_words = 0;
    for (DevState &d : ctx->devs) {
        if (!d.ran) return fail(ctx, LF_ERR_BAD_ARG, "download before run");
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        if (d.n_tasks == 0) continue;
        if (d.ops_words) {
            if (!ops || (base_words + d.ops_words) * 4 > ops_cap) return fail(ctx, LF_ERR_OPS_CAPACITY, "ops buffer too small");
            LF_TRY(lfb_d2h(ops + base_words * 4, d.ops.p, (size_t)d.ops_words * 4, d.stream));
        }
        LF_TRY(lfb_d2h(res + d.task_first, d.res.p, (size_t)d.n_tasks * sizeof(lf_align_result), d.stream));
        base_words += d.ops_words;
    }
    base_words = 0;
    for (DevState &d : ctx->devs) {
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        LF_TRY(lfb_sync(d.stream));
        if (base_words) for (size_t i = 0; i < d.n_tasks; i++) res[d.task_first + i].ops_off += base_words * 16;
        base_words += d.ops_words;
    }
    return LF_OK;
}

int lf_gpu_align_batch(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_align_task *tasks, size_t n,
                       lf_align_result *res, uint8_t *ops, size_t ops_cap)
{
    int rc;
    if (n == 0) return LF_OK;
    if ((rc = lf_gpu_upload_reads(ctx, reads)) != 0) return rc;
    if ((rc = lf_gpu_upload_align_tasks(ctx, tasks, n)) != 0) return rc;
    if ((rc = lf_gpu_run_align(ctx)) != 0) return rc;
    if ((rc = lf_gpu_download_align(ctx, res, ops, ops_cap)) != 0) return rc;
    return lf_gpu_sync(ctx);
}

/* ---- extension ---- */
int lf_gpu_upload_extend_tasks(lf_gpu_ctx *ctx, const lf_extend_task *tasks, size_t n)
{
    if (!ctx || (!tasks && n) || n > 0x7fffffffu) return LF_ERR_BAD_ARG;
    const size_t nd = ctx->devs.size();
    for (size_t k = 0; k < nd; k++) {
        DevState &d = ctx->devs[k];
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        const size_t lo = n * k / nd, hi = n * (k + 1) / nd;
        d.etask_first = lo; d.n_etasks = (uint32_t)(hi - lo);
        LF_TRY(d.etasks.reserve((hi - lo) * sizeof(lf_extend_task)));
        if (hi > lo) LF_TRY(lfb_h2d(d.etasks.p, tasks + lo, (hi - lo) * sizeof(lf_extend_task), d.stream));
    }
    return LF_OK;
}

static int run_extend_dev(lf_gpu_ctx *ctx, DevState &d, lfb_stream s, size_t scr_bound = 0)
{   /* scr_bound != 0: the caller's bound on the scratch items (sum of what k_extend_prep asks for): no host round trip */
    const size_t n = d.n_etasks;
    if (!n) return LF_OK;
    LF_TRY(d.eres.reserve(n * sizeof(lf_extend_result)));
    LF_TRY(d.escr_items.reserve(n * 4)); LF_TRY(d.escr_off.reserve((n + 1) * 8));
    LfExtDev v;
    v.pac = d.pac.as<uint8_t>(); v.l_pac = ctx->l_pac; v.bases = d.bases.as<uint8_t>(); v.read_off = d.read_off.as<uint64_t>(); v.n_reads = d.n_reads;
    v.tasks = d.etasks.as<lf_extend_task>(); v.n_tasks = (uint32_t)n; v.res = d.eres.as<lf_extend_result>();
    v.scr_off = d.escr_off.as<uint64_t>(); v.scratch = nullptr;
    LFB_LAUNCH(k_extend_prep, (unsigned)((n + 255) / 256), 256, 0, s, v, d.escr_items.as<uint32_t>());
    LF_TRY(lfb_scan_excl_total(d.tmp, d.escr_items.as<uint32_t>(), d.escr_off.as<unsigned long long>(), n, s));
    if (scr_bound) LF_TRY(d.escr.reserve(scr_bound * sizeof(int2) + 64));
    else {
        HostTotals *ht = (HostTotals *)d.pinned;
        LF_TRY(lfb_d2h(&ht->scr_total, d.escr_off.as<unsigned long long>() + n, 8, s));
        LF_TRY(lfb_sync(s));
        LF_TRY(d.escr.reserve((size_t)ht->scr_total * sizeof(int2) + 64));
    }
    v.scratch = d.escr.as<int2>();
    LFB_LAUNCH(k_ksw_extend, (unsigned)((n + 3) / 4), 128, 0, s, v); /* one warp per task */
    LF_TRY(lfb_last_error());
    ctx->stats.extend_tasks += n;
    return LF_OK;
}

int lf_gpu_run_extend(lf_gpu_ctx *ctx)
{
    if (!ctx) return LF_ERR_BAD_ARG;
    for (DevState &d : ctx->devs) {
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        int rc = run_extend_dev(ctx, d, d.stream);
        if (rc) return rc;
    }
    ctx->stats.kernel_launches = lfb_launches;
    return LF_OK;
}

/* Chain operator, single-device contexts: extensions on their own stream, started once everything enqueued on the main
 * stream so far (reads, round-1 tasks) has arrived, running beside the round-1 kernels, without a host round trip;
 * results go to `out` (pinned) asynchronously -- spec_extend_wait() before reading them. */
static int spec_extend_start(lf_gpu_ctx *ctx, const lf_extend_task *tasks, size_t n, lf_extend_result *out)
{
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    d.etask_first = 0; d.n_etasks = (uint32_t)n;
    if (!n) return LF_OK;
    LF_TRY(d.etasks.reserve(n * sizeof(lf_extend_task) + 64));
    /* no wait on the main stream: the caller starts this after lf_gpu_run_align, whose class-count sync has already
     * waited for the reads and tasks to arrive, and waiting now would put the extensions behind the round-1 kernels */
    LF_TRY(h2d_k(d, d.etasks.p, tasks, n * sizeof(lf_extend_task), d.ext_stream, true));   /* `tasks` is the caller's pinned staging */
    size_t bound = 0;
    for (size_t i = 0; i < n; i++) bound += (size_t)tasks[i].q_len + 1u + ((size_t)tasks[i].q_len + 7u) / 8u + 1u;   /* k_extend_prep's figure */
    int rc = run_extend_dev(ctx, d, d.ext_stream, bound);
    if (rc) return rc;
    LF_TRY(lfb_d2h(out, d.eres.p, n * sizeof(lf_extend_result), d.ext_stream));
    return LF_OK;
}
static int spec_extend_wait(lf_gpu_ctx *ctx)
{
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    LF_TRY(lfb_sync(d.ext_stream));
    return LF_OK;
}

int lf_gpu_download_extend(lf_gpu_ctx *ctx, lf_extend_result *res)
{
    if (!ctx || !res) return LF_ERR_BAD_ARG;
    for (DevState &d : ctx->devs) {
        if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
        if (d.n_etasks) LF_TRY(lfb_d2h(res + d.etask_first, d.eres.p, (size_t)d.n_etasks * sizeof(lf_extend_result), d.stream));
    }
    return lf_gpu_sync(ctx);
}

int lf_gpu_extend_batch(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_extend_task *tasks, size_t n, lf_extend_result *res)
{
    int rc;
    if (n == 0) return LF_OK;
    if (reads && (rc = lf_gpu_upload_reads(ctx, reads)) != 0) return rc; /* reads == NULL: keep the resident chunk */
    if ((rc = lf_gpu_upload_extend_tasks(ctx, tasks, n)) != 0) return rc;
    if ((rc = lf_gpu_run_extend(ctx)) != 0) return rc;
    return lf_gpu_download_extend(ctx, res);
}

int lf_gpu_get_stats(const lf_gpu_ctx *ctx, lf_gpu_stats *out)
{
    if (!ctx || !out) return LF_ERR_BAD_ARG;
    *out = ctx->stats;
    out->kernel_launches = lfb_launches;
    return LF_OK;
}

int lf_gpu_class_timeline(lf_gpu_ctx *ctx, float *start_ms, float *end_ms, int n)
{ /* start / end of each size-class kernel of the last run on device 0, relative to the first launch; -1 = not run */
    if (!ctx || !start_ms || !end_ms || n < LF_NCLS) return LF_ERR_BAD_ARG;
    DevState &d = ctx->devs[0];
    for (int c = 0; c < n; c++) { start_ms[c] = -1.f; end_ms[c] = -1.f; }
#ifndef LF_EMU
    if (!d.ran) return LF_OK;
    for (int c = 0; c < LF_NCLS; c++) {
        if (!d.cls_ran[c]) continue;
        cudaEventElapsedTime(&start_ms[c], d.ev[2], d.cls_ev[c][0]);
        cudaEventElapsedTime(&end_ms[c], d.ev[2], d.cls_ev[c][1]);
    }
#endif
    return LF_OK;
}

int lf_gpu_class_counts(lf_gpu_ctx *ctx, uint32_t *counts, int n)
{
    if (!ctx || !counts || n < LF_NCLS) return LF_ERR_BAD_ARG;
    for (int c = 0; c < n; c++) counts[c] = c < LF_NCLS ? ctx->devs[0].cls_count[c] : 0u;
    return LF_OK;
}

int lf_gpu_int32_peak(lf_gpu_ctx *ctx, int which, double *tops)
{
    if (!ctx || !tops || which < 0 || which > 3) return LF_ERR_BAD_ARG;
    DevState &d = ctx->devs[0];
    if (set_dev(d)) return fail(ctx, LF_ERR_CUDA, "cudaSetDevice");
    const int grid = 148 * 8, block = 256, iters = 4096;
    LF_TRY(d.queue.reserve(64));
    uint32_t *sink = d.queue.as<uint32_t>();
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
#ifndef LF_EMU
        cudaEventRecord(d.ev[0], d.stream);
#endif
        switch (which) {
        case 0: { auto kern = k_int32_peak<0>; LFB_LAUNCH(kern, grid, block, 0, d.stream, sink, iters, 12345u + rep); } break;
        case 1: { auto kern = k_int32_peak<1>; LFB_LAUNCH(kern, grid, block, 0, d.stream, sink, iters, 12345u + rep); } break;
        case 2: { auto kern = k_int32_peak<2>; LFB_LAUNCH(kern, grid, block, 0, d.stream, sink, iters, 12345u + rep); } break;
        default: { auto kern = k_int32_peak<3>; LFB_LAUNCH(kern, grid, block, 0, d.stream, sink, iters, 12345u + rep); } break;
        }
#ifndef LF_EMU
        cudaEventRecord(d.ev[1], d.stream);
        LF_TRY(lfb_sync(d.stream));
        float ms = 0; cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]);
        if (rep && ms < best_ms) best_ms = ms;
#else
        best_ms = 1.0f;
#endif
    }
    *tops = (double)grid * block * (double)iters * 64.0 / (best_ms * 1e-3) / 1e12;
    return LF_OK;
}

} /* extern "C" */

#include "lf_chain.inl"
#include "lf_seed.inl"
