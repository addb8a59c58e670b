/*
 * lf_oracle.h -- CPU restatement of lordFAST's per-candidate alignment stage.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared with; it is
 * never linked into liblfgpu.so and never used as a fallback.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * It restates, from the behavioural rules in SURVEY.md Appendix A-C, what these reference
 * functions compute (file:line relative to /root/reference):
 *   edlibAlign NW/SHW + path      lib/edlib/edlib.cpp:101-221, 475-629, 657-858,
 *                                 872-1071 (traceback), 1090-1143 (size rule),
 *                                 1161-1330 (Hirschberg split rule)
 *   ksw_extend2                   lib/bwa/ksw.c:380-479
 *   bwt_str_pac2char/_get_pac     src/BWT.cpp:310, 593-607
 *   reverseComplement             src/Common.cpp:34-66
 *   convertChar2int / rcIntStr    src/LordFAST.cpp:1191-1201
 *   alignChain_edlib              src/LordFAST.cpp:1765-2258 (+ CIGAR/MD helpers 1570-1763)
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's own code compiled into oracle/_ref
 * (tests/test_oracle_vs_ref.py) and against fixtures generated from it (tests/golden/).
 */
#ifndef LF_ORACLE_H
#define LF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { LFO_MODE_NW = 0, LFO_MODE_SHW = 1 };

typedef struct {
    int edit_distance; /* Levenshtein distance (NW) or min over target prefixes (SHW) */
    int end_location;  /* edlib endLocations[0]: tlen-1 for NW; may be -1 for SHW      */
    int n_ops;         /* alignmentLength (0 if want_path == 0)                         */
} lfo_align_out;

/* ops: edlib alignment codes 0=match 1=insert(query only) 2=delete(target only) 3=mismatch,
 * capacity >= qlen + tlen.  Sequences are compared as raw bytes.  Returns 0 on success. */
int lfo_align(const char *q, int qlen, const char *t, int tlen, int mode, int want_path,
              lfo_align_out *out, uint8_t *ops);

/* ksw_extend2 semantics; returns the best score, *qle / *tle as the reference. */
int lfo_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m,
                    const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w,
                    int end_bonus, int zdrop, int h0, int *qle, int *tle);

/* 2-bit reference helpers (bwa .pac layout: base l in byte l>>2, bits ((~l)&3)<<1). */
int  lfo_pac_get(const uint8_t *pac, uint32_t l);
void lfo_pac2char(const uint8_t *pac, uint32_t beg, uint32_t len, char *dst);
void lfo_pac2int(const uint8_t *pac, uint32_t beg, uint32_t len, uint8_t *dst);
void lfo_pack_ref(const char *seq, int64_t len, uint8_t *pac); /* ACGT only */
void lfo_revcomp(const char *src, char *dst, int len);         /* dst gets len+1 bytes */

/* One SAM-level alignment record produced by a chain (reference Sam_t, src/LordFAST.h:81-100). */
typedef struct {
    uint32_t flag, pos, posEnd, qStart, qEnd;
    int32_t  nmCount;
    char    *cigar; /* malloc'd */
    char    *md;    /* malloc'd */
} lfo_sam;

typedef struct {
    uint32_t tPos, qPos, len; /* reference Seed_t (bit-fields unpacked) */
} lfo_seed;

typedef struct {
    const uint8_t *pac;
    int64_t        l_pac;
    int            n_contigs;
    const int64_t *contig_off;
    const int32_t *contig_len;
} lfo_ref;

/* Counters describing the work one chain issued (for benches / fixtures). */
typedef struct {
    int64_t n_align, n_extend, cells_align, cells_extend;
} lfo_stats;

/* alignChain_edlib: `query` is the oriented read (read or its reverse complement), n_seeds >= 2.
 * Appends up to *n_sam records (caller frees cigar/md with lfo_free_sam).  Returns 0. */
int  lfo_align_chain(const lfo_ref *ref, const lfo_seed *seeds, int n_seeds, const char *query,
                     int32_t read_len, int is_rev, lfo_sam *sam, int sam_cap, int *n_sam,
                     lfo_stats *stats);
void lfo_free_sam(lfo_sam *sam, int n);

#ifdef __cplusplus
}
#endif
#endif
