/*
 * lf_gpu.h -- C ABI of liblfgpu.so: lordFAST's per-candidate alignment stage as one batched
 * B200 (sm_100a) path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * What each entry point replaces in the reference (file:line relative to vpc-ccg/lordfast):
 *
 *   lf_gpu_align_batch    every edlibAlign(.., edlibNewAlignConfig(-1, NW|SHW, PATH)) call that
 *                         alignChain_edlib issues: src/LordFAST.cpp:1833 (head, SHW), :1941 (gap,
 *                         NW), :2168 (tail, SHW) and the follow-ups :1853, :2001, :2037, :2039,
 *                         :2084, :2184; library seam lib/edlib/edlib.h:190-192.  The reference
 *                         fetch of the target slice (bwt_str_pac2char, src/BWT.cpp:601-607),
 *                         reverseComplement (src/Common.cpp:57-66) and edlib's transformSequences /
 *                         buildPeq (lib/edlib/edlib.cpp:281, :1350) are fused into the kernels.
 *   lf_gpu_extend_batch   ksw_extend (:1848, :2180) and ksw_extend2 (:1971, :1981); library seam
 *                         lib/bwa/ksw.h:107-108; convertChar2int / reverseComplementIntStr /
 *                         bwt_str_pac2int (src/LordFAST.cpp:1191-1201, src/BWT.cpp:593-599) fused.
 *   lf_gpu_init           the in-memory 2-bit reference `_fmd_index->pac` (lib/bwa/bwa.c:270-275,
 *                         layout lib/bwa/bntseq.c:224-225) is copied to HBM once.
 *
 * Results are bit-identical to the reference calls: editDistance, endLocations[0] and the
 * alignment[] op codes (lib/edlib/edlib.h:69-72, :121-166); score / qle / tle of ksw_extend2.
 *
 * Conventions: every function returns LF_OK (0) or a negative lf_status and never exits the
 * process; all buffers are caller-owned; results come back in task order; a context is bound to
 * its device set; distinct contexts may be used concurrently from different host threads, one
 * context must not be.  There is no CPU fallback: without a CUDA device lf_gpu_init fails with
 * LF_ERR_NO_DEVICE.
 */
#ifndef LF_GPU_H
#define LF_GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lf_gpu_ctx lf_gpu_ctx;

typedef enum {
    LF_OK = 0,
    LF_ERR_NO_DEVICE = -1,   /* no usable CUDA device / driver */
    LF_ERR_CUDA = -2,        /* a CUDA call failed; see lf_gpu_last_error */
    LF_ERR_BAD_ARG = -3,     /* null pointer, zero-length sequence, offset out of range ... */
    LF_ERR_OPS_CAPACITY = -4,/* `ops` buffer smaller than lf_gpu_ops_capacity() asks for */
    LF_ERR_NOMEM = -5
} lf_status;

/* ---- inputs ------------------------------------------------------------------------------- */

/* The reads of one chunk as sequenced (forward); the GPU derives reverse complements itself.
 * bases[offsets[r] .. offsets[r+1]) is read r (ASCII; anything but upper-case ACGT never matches
 * the reference in the edit-distance kernels, exactly as in edlib's byte comparison). */
typedef struct {
    const uint8_t  *bases;
    const uint64_t *offsets; /* n_reads + 1 entries */
    uint32_t        n_reads;
} lf_reads;

enum { LF_MODE_NW = 0, LF_MODE_SHW = 1 }; /* EDLIB_MODE_NW / EDLIB_MODE_SHW */

enum {
    LF_F_READ_REV     = 1, /* oriented read = reverse complement of the stored read (isRev)        */
    LF_F_REVERSE_BOTH = 2, /* query and target slices are both taken right-to-left: read head
                              (:1827-1831), head re-alignment (:1853), second split part (:2082-2084).
                              The reference also complements both sides, which leaves equality as is. */
    LF_F_RC_QUERY     = 4, /* only the query slice is reverse-complemented (inversion test :2038-2039) */
    LF_F_NO_PATH      = 8  /* distance / end location only (EDLIB_TASK_DISTANCE)                    */
};

typedef struct {        /* 24 bytes */
    uint32_t read_id;   /* index into lf_reads                                                     */
    uint32_t q_off;     /* slice start in the ORIENTED read (the read, or its reverse complement) */
    uint32_t q_len;     /* >= 1                                                                    */
    uint32_t t_off;     /* slice start on the forward reference (pac coordinate)                   */
    uint32_t t_len;     /* >= 1                                                                    */
    uint16_t flags;     /* LF_F_*                                                                  */
    uint8_t  mode;      /* LF_MODE_*                                                               */
    uint8_t  reserved;
} lf_align_task;

typedef struct {        /* 24 bytes */
    int32_t  edit_distance; /* EdlibAlignResult.editDistance                                       */
    int32_t  end_location;  /* endLocations[0]: t_len-1 for NW; first best prefix end for SHW, may be -1 */
    uint64_t ops_off;       /* first op of this task in the 2-bit op stream (op index, not bytes)  */
    uint32_t ops_len;       /* alignmentLength                                                     */
    int32_t  status;        /* 0, or a negative lf_status for this task                            */
} lf_align_result;

/* Op stream: 2 bits per op, four ops per byte, op p at bits 2*(p&3) of byte p>>2; codes are
 * edlib's 0 = match, 1 = insert (query base only), 2 = delete (target base only), 3 = mismatch,
 * in the order edlib returns them (left to right in the task's own orientation). */
#define LF_OP_AT(ops, p) ((unsigned)(((const uint8_t *)(ops))[(p) >> 2] >> (((p) & 3) << 1)) & 3u)

typedef struct {        /* 52 bytes; sequences described as in lf_align_task */
    uint32_t read_id, q_off, q_len, t_off, t_len;
    uint16_t flags;     /* LF_F_READ_REV, LF_F_REVERSE_BOTH                                        */
    uint8_t  matrix;    /* LF_MAT_* */
    uint8_t  reserved;
    int32_t  o_del, e_del, o_ins, e_ins, w, zdrop, h0; /* ksw_extend2 arguments; end_bonus = 0     */
} lf_extend_task;

enum { LF_MAT_CLIP = 0 /* _pf_kswMatrix_clip: +2 / -16, N row/col 0 (src/LordFAST.cpp:82-83, 178-187) */,
       LF_MAT_DEFAULT = 1 /* _pf_kswMatrix: +2 / -5 (:78-79, 166-176) */ };

typedef struct {        /* 12 bytes */
    int32_t score, qle, tle; /* return value, *_qle, *_tle of ksw_extend2                          */
} lf_extend_result;

/* ---- life cycle --------------------------------------------------------------------------- */

/* pac: l_pac/4+1 bytes in bwa's layout (base l in byte l>>2, bits ((~l)&3)<<1).  devices == NULL
 * or n_devices == 0 selects the current CUDA device.  The reference is replicated on every
 * device of the set and a batch is split across them by contiguous task ranges. */
int  lf_gpu_init(lf_gpu_ctx **ctx, const uint8_t *pac, int64_t l_pac, const int *devices, int n_devices);
void lf_gpu_destroy(lf_gpu_ctx *ctx);
const char *lf_gpu_last_error(const lf_gpu_ctx *ctx);

/* Starts CUDA (driver, context, module load: seconds on a cold box) on a background thread and returns at once, so
 * that it overlaps the caller's own start-up -- lordFAST calls it first thing and then loads its index
 * (bwt_load, src/baseFAST.cpp:56); lf_gpu_init joins it.  Optional; safe to call more than once. */
void lf_gpu_prewarm(void);

/* Pinned host memory for task / result / op buffers (pageable memory works, pinned is faster). */
void *lf_gpu_host_alloc(size_t bytes);
void  lf_gpu_host_free(void *p);

/* Bytes the caller must provide in `ops` for these tasks (a slot of q_len+t_len ops per task,
 * rounded up to 16 ops). */
size_t lf_gpu_ops_capacity(const lf_align_task *tasks, size_t n);

/* ---- the hot path ------------------------------------------------------------------------- */

/* Host buffers in, host buffers out: copies reads/tasks to the device(s), runs the alignment
 * kernels, copies results and ops back.  `ops` may be NULL if every task has LF_F_NO_PATH. */
int lf_gpu_align_batch(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_align_task *tasks, size_t n,
                       lf_align_result *res, uint8_t *ops, size_t ops_cap);

int lf_gpu_extend_batch(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_extend_task *tasks, size_t n,
                        lf_extend_result *res);

/* The same path split into its three phases, so that a caller (and bench.py) can keep a batch
 * resident in HBM: upload -> run (any number of times) -> download. */
int lf_gpu_upload_reads(lf_gpu_ctx *ctx, const lf_reads *reads);
int lf_gpu_pack_reads(lf_gpu_ctx *ctx);   /* rebuilds the bit planes of the resident reads (k_pack_reads; part of every upload, callable on its own for timing) */
int lf_gpu_upload_align_tasks(lf_gpu_ctx *ctx, const lf_align_task *tasks, size_t n);
int lf_gpu_run_align(lf_gpu_ctx *ctx);  /* kernels only, on the resident batch; asynchronous */
int lf_gpu_sync(lf_gpu_ctx *ctx);
int lf_gpu_download_align(lf_gpu_ctx *ctx, lf_align_result *res, uint8_t *ops, size_t ops_cap);
int lf_gpu_upload_extend_tasks(lf_gpu_ctx *ctx, const lf_extend_task *tasks, size_t n);
int lf_gpu_run_extend(lf_gpu_ctx *ctx);
int lf_gpu_download_extend(lf_gpu_ctx *ctx, lf_extend_result *res);

/* ---- the chain-level operator -------------------------------------------------------------- */

/* alignChain_edlib (src/LordFAST.cpp:1765-2258, reached through the reference's hook
 * `void (*alignChain)(Chain_t&, char*, int32_t, int, SamList_t&)`, :107) for a whole chunk of
 * candidate chains: round-1 alignments (the task list is derived from the chains on the device), the
 * clip / split triggers (the reference's float expressions, evaluated on the device; the split /
 * inversion decisions that need follow-up distances on the host), ksw extensions, follow-up
 * alignments, and the reference's CIGAR / MD accumulation (on the device for single-device contexts).  One lf_sam_record per Sam_t the reference would push to map.samList,
 * in chain order.  The caller (lordFAST's alignWin, :1063-1083) computes alnScore / totalScore from
 * nmCount, qStart, qEnd exactly as today. */
typedef struct { uint32_t tPos, qPos, len; } lf_seed;      /* Seed_t (src/LordFAST.h:30-35), bit-fields unpacked */
typedef struct {                                          /* Chain_t + the call's query / isRev arguments */
    uint64_t seed_off;   /* first seed in the seeds array                                           */
    uint32_t n_seeds;    /* chainLen, >= 2 (alignWin only calls alignChain then, :1063)              */
    uint32_t read_id;    /* index into lf_reads                                                     */
    uint32_t is_rev;     /* 1: query = reverse complement of the stored read                        */
    uint32_t reserved;
} lf_chain;
typedef struct { const int64_t *offset; const int32_t *len; int32_t n; } lf_contigs; /* bns->anns[] offsets / lengths */
typedef struct {         /* Sam_t fields alignChain_edlib fills (src/LordFAST.h:81-100) */
    uint32_t chain_id, flag, pos, posEnd, qStart, qEnd;
    int32_t  nmCount;
    uint32_t cigar_len, md_len;
    uint64_t cigar_off, md_off; /* NUL-terminated strings inside lf_chain_results_text() */
} lf_sam_record;
typedef struct {
    uint64_t round1_tasks, round2_extends, round3_tasks, records;
    float ms_tasks, ms_round1, ms_rounds23, ms_emit, ms_merge; /* host wall time of the phases of the call */
    float reserved;
} lf_chain_stats;
typedef struct lf_chain_results lf_chain_results;         /* library-owned, free with lf_chain_results_free */

/* pac_host: the same 2-bit reference that was given to lf_gpu_init (MD strings need reference bases).
 * reads->bases == NULL (offsets and n_reads still given): the reads are the ones the last lf_gpu_upload_reads /
 * lf_gpu_seed_batch left on the device -- no second trip over PCIe for a chunk that was just seeded (single device). */
int lf_gpu_align_chains(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_contigs *contigs, const lf_seed *seeds,
                        const lf_chain *chains, size_t n_chains, const uint8_t *pac_host, lf_chain_results **out);
const lf_sam_record *lf_chain_results_records(const lf_chain_results *r, size_t *n);
const char *lf_chain_results_text(const lf_chain_results *r, size_t *bytes);
int  lf_chain_results_stats(const lf_chain_results *r, lf_chain_stats *out);
void lf_chain_results_free(lf_chain_results *r);

/* ---- FM-index seeding ---------------------------------------------- */
/* getLocs_extend_whole_step (src/BWT.cpp:312-394), the seeding step mapSeq runs per read (src/LordFAST.cpp:507), for a
 * batch of reads: SAMPLING_COUNT sample positions per read, at each the longest exact match of at least MIN_ANCHOR_LEN
 * bases against bwa's FM index of reference + reverse complement (bwt_count_exact_cached, src/BWT.cpp:265-298, on
 * lib/bwa/bwt.c:86-143), kept if it has fewer than MAX_REF_HITS occurrences and is not contained in the previous kept
 * match, every occurrence located through the sampled suffix array (bwt_sa, lib/bwa/bwt.c:86-98) and appended to the
 * forward or to the reverse seed list of the read, in the reference's order. */
typedef struct { uint64_t beg, end; } lf_fm_cache_entry;     /* bwtCache_t, src/BWT.h:41-46 */
typedef struct {                /* the members of bwa's bwt_t (lib/bwa/bwt.h:44-57) that seeding reads, and bns->l_pac */
    const uint32_t *bwt;        /* bwt_t::bwt: per 128 bases four 64-bit occurrence counts, then eight words of 16 bases */
    uint64_t bwt_size;          /* 32-bit words in bwt */
    uint64_t primary, L2[5], seq_len;
    const uint64_t *sa;         /* bwt_t::sa, n_sa entries, sa[0] = (uint64_t)-1 */
    uint64_t n_sa;
    int32_t  sa_intv;           /* power of two */
    int32_t  k_cache;           /* kCache (src/BWT.cpp:34): 12 in the reference; <= min_anchor_len */
    const lf_fm_cache_entry *cache; /* _fmd_cacheTable, 4^k_cache entries; NULL: derived from bwt on the device (the
                                       table bwt_cache_gen would write, src/BWT.cpp:60-138) */
    int64_t  l_pac;             /* bns->l_pac: suffix positions >= l_pac are on the reverse strand */
} lf_fm_index;
typedef struct { int32_t min_anchor_len, sampling_count, max_ref_hits; } lf_seed_params; /* MIN_ANCHOR_LEN, SAMPLING_COUNT, MAX_REF_HITS */
typedef struct lf_seed_results lf_seed_results;   /* library-owned; valid until the next lf_gpu_seed_batch on the context or lf_seed_results_free */

/* Copies the index to device 0 of the context (and builds the k-mer table there if fm->cache is NULL). */
int lf_gpu_seed_init(lf_gpu_ctx *ctx, const lf_fm_index *fm);
/* reads == NULL: the reads of the last lf_gpu_upload_reads / lf_gpu_seed_batch are still resident and are used again.
 * Bases other than ACGTacgt end a match, as does the end of the read (the reference reads the NUL there). */
int lf_gpu_seed_batch(lf_gpu_ctx *ctx, const lf_reads *reads, const lf_seed_params *params, lf_seed_results **out);
/* reverse = 0: seedForward lists, 1: seedReverse lists.  offsets: n_reads + 1 entries into the returned array.
 * qPos and len carry Seed_t's bit-field widths (20 and 12 bits, src/LordFAST.h:30-35). */
const lf_seed *lf_seed_results_list(const lf_seed_results *r, int reverse, const uint64_t **offsets, size_t *n);
void lf_seed_results_free(lf_seed_results *r);
/* The k-mer table in use on the device (4^k_cache entries), for checks against _fmd_cacheTable. */
int lf_gpu_seed_cache_download(lf_gpu_ctx *ctx, lf_fm_cache_entry *out, size_t n);
/* Work and CUDA-event times of the last lf_gpu_seed_batch on the context. */
typedef struct {
    float    search_ms;      /* positions + longest-match search + containment filter + hit offsets */
    float    locate_ms;      /* locate + strand split */
    uint64_t positions;      /* n_reads * sampling_count */
    uint64_t hits;           /* seeds written (both lists) */
    uint64_t search_steps;   /* backward-search steps taken: two occurrence lookups (one or two 64-byte bwt blocks) each */
    uint64_t locate_steps;   /* inverse-psi steps taken by bwt_sa: one bwt block each */
} lf_seed_stats;
int lf_gpu_seed_stats(lf_gpu_ctx *ctx, lf_seed_stats *out);

/* ---- measurement hooks -------------------------------------------------------------------- */

typedef struct {
    uint64_t kernel_launches;   /* launches of our own kernels since the context was created       */
    uint64_t align_tasks, extend_tasks;
    uint64_t cells;             /* sum of q_len * t_len over align tasks                           */
    uint64_t word_columns;      /* sum of ceil(q_len/32) * t_len: 16 INT32 ops each    */
    float    last_run_ms;       /* CUDA-event time of the last lf_gpu_run_* on device 0's stream   */
    float    last_main_kernel_ms; /* CUDA-event time of the dominant (small-task) kernels therein  */
    uint64_t last_main_word_columns;
} lf_gpu_stats;
int lf_gpu_get_stats(const lf_gpu_ctx *ctx, lf_gpu_stats *out);

/* Timeline of the last lf_gpu_run_align on device 0: start / end (ms after the first launch) of each size-class
 * kernel; index 2*i+shw for the register classes (NW = 1,2,3,4,6,8,12,16), 16 = large-task kernel, 18..21 = the
 * banded register classes (near-diagonal global tasks of NW = 6,8,12,16), 22..28 = the lane-group classes; -1 = not run.
 * n >= 29. */
int lf_gpu_class_timeline(lf_gpu_ctx *ctx, float *start_ms, float *end_ms, int n);

/* Tasks per size class of the last lf_gpu_run_align on device 0 (same indices as lf_gpu_class_timeline; 22..28 = the
 * lane-group classes of k_myers_group: path 9-16 / 17-32 / 33-64 words, distance-only 17-32 / 33-64 / 65-128 / 129-256 words). */
int lf_gpu_class_counts(lf_gpu_ctx *ctx, uint32_t *counts, int n);

/* INT32 issue-rate microbenchmarks on device 0 (dependent-free streams on all SMs); Top/s.
 * which: 0 = LOP3 only, 1 = IADD3 only, 2 = LOP3+IADD3 mix, 3 = LOP3 + IMAD mix */
int lf_gpu_int32_peak(lf_gpu_ctx *ctx, int which, double *tops);

#ifdef __cplusplus
}
#endif
#endif
