/*
 * lf_oracle.c -- CPU restatement of lordFAST's per-candidate alignment stage (see lf_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path, never a fallback for it.
 *
 * The edit-distance part is written from the definition (plain Levenshtein DP, no bit-vectors,
 * no band): SURVEY.md Appendix A shows that everything lordFAST consumes from edlibAlign is a
 * band-independent pure function of the two strings.  tests/test_oracle_vs_ref.py pins this
 * against the reference's own edlib.cpp / ksw.c compiled into oracle/_ref.
 */
#include "lf_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------------------------------ */
/* Levenshtein columns                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* Final column D(0..qlen, tlen) of the unit-cost DP.  `rev` walks both strings backwards
 * (the "right half" pass of the split rule, lib/edlib/edlib.cpp:1190-1196).
 * If last_row != NULL it receives D(qlen, j) for j = 0..tlen (SHW scan, edlib.cpp:583-618). */
static void lev_column(const char *q, int qlen, const char *t, int tlen, int rev, int *col,
                       int *last_row)
{
    int i, j;
    for (i = 0; i <= qlen; i++) col[i] = i; /* D(i,0) = i */
    if (last_row) last_row[0] = qlen;
    for (j = 1; j <= tlen; j++) {
        char tc = rev ? t[tlen - j] : t[j - 1];
        int diag = col[0]; /* D(i-1, j-1) */
        col[0] = j;        /* D(0,j) = j */
        for (i = 1; i <= qlen; i++) {
            char qc = rev ? q[qlen - i] : q[i - 1];
            int up = col[i - 1], left = col[i];
            int best = diag + (qc != tc);
            if (up + 1 < best) best = up + 1;
            if (left + 1 < best) best = left + 1;
            diag = left;
            col[i] = best;
        }
        if (last_row) last_row[j] = col[qlen];
    }
}

typedef struct {
    uint8_t *p;
    int n;
} opbuf;

static void emit(opbuf *o, int code, int count)
{
    if (o->p) memset(o->p + o->n, code, (size_t)count);
    o->n += count;
}

/* Leaf: full matrix + canonical walk from the bottom-right corner, preferring
 * Up (op 1) over Left (op 2) over Diagonal (op 0 / 3)  -- edlib.cpp:950, 984, 1015-1016. */
static int leaf_traceback(const char *q, int qlen, const char *t, int tlen, opbuf *o)
{
    size_t W = (size_t)tlen + 1;
    int *D = (int *)malloc(sizeof(int) * (size_t)(qlen + 1) * W);
    uint8_t *tmp;
    int i, j, n = 0;
    if (!D) return -1;
    for (j = 0; j <= tlen; j++) D[j] = j;
    for (i = 1; i <= qlen; i++) {
        int *row = D + (size_t)i * W, *prev = row - W;
        row[0] = i;
        for (j = 1; j <= tlen; j++) {
            int best = prev[j - 1] + (q[i - 1] != t[j - 1]);
            if (prev[j] + 1 < best) best = prev[j] + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            row[j] = best;
        }
    }
    tmp = (uint8_t *)malloc((size_t)qlen + tlen + 1);
    if (!tmp) { free(D); return -1; }
    i = qlen; j = tlen;
    while (i > 0 || j > 0) {
        int cur = D[(size_t)i * W + j];
        if (i > 0 && D[(size_t)(i - 1) * W + j] + 1 == cur) { tmp[n++] = 1; i--; }
        else if (j > 0 && D[(size_t)i * W + j - 1] + 1 == cur) { tmp[n++] = 2; j--; }
        else { tmp[n++] = (q[i - 1] == t[j - 1]) ? 0 : 3; i--; j--; }
    }
    if (o->p) for (i = 0; i < n; i++) o->p[o->n + i] = tmp[n - 1 - i];
    o->n += n;
    free(tmp); free(D);
    return 0;
}

/* obtainAlignment (edlib.cpp:1090-1143) + obtainAlignmentHirschberg (1161-1330). */
static int path_rec(const char *q, int qlen, const char *t, int tlen, int best, opbuf *o)
{
    long long blocks64, est;
    if (qlen == 0) { emit(o, 2, tlen); return 0; }
    if (tlen == 0) { emit(o, 1, qlen); return 0; }
    blocks64 = (qlen + 63) / 64; /* the rule counts 64-bit blocks whatever word the solver uses */
    est = 20LL * blocks64 * tlen + 8LL * tlen;
    if (est < 1024 * 1024 || tlen < 2) return leaf_traceback(q, qlen, t, tlen, o);
    {
        int lw = tlen / 2, rw = tlen - lw, x, found = -1, ls = 0, rs = 0, rc;
        int *L = (int *)malloc(sizeof(int) * ((size_t)qlen + 1));
        int *R = (int *)malloc(sizeof(int) * ((size_t)qlen + 1));
        if (!L || !R) { free(L); free(R); return -1; }
        lev_column(q, qlen, t, lw, 0, L, NULL);      /* L[x] = D(x, lw)                       */
        lev_column(q, qlen, t + lw, rw, 1, R, NULL); /* R[y] = dist(last y of q, right half) */
        /* smallest interior row first (edlib.cpp:1263-1271), then the top boundary
         * (1273-1280), then the bottom one (1281-1289) */
        for (x = 1; x <= qlen - 1; x++)
            if (L[x] + R[qlen - x] == best) { found = x; ls = L[x]; rs = R[qlen - x]; break; }
        if (found < 0 && lw + R[qlen] == best) { found = 0; ls = lw; rs = R[qlen]; }
        if (found < 0 && L[qlen] + rw == best) { found = qlen; ls = L[qlen]; rs = rw; }
        free(L); free(R);
        if (found < 0) return -2;
        rc = path_rec(q, found, t, lw, ls, o);
        if (rc) return rc;
        return path_rec(q + found, qlen - found, t + lw, rw, rs, o);
    }
}

int lfo_align(const char *q, int qlen, const char *t, int tlen, int mode, int want_path,
              lfo_align_out *out, uint8_t *ops)
{
    int *col, *row = NULL, ed, end;
    opbuf o;
    if (qlen <= 0 || tlen <= 0) return -1; /* lordFAST never issues these */
    col = (int *)malloc(sizeof(int) * ((size_t)qlen + 1));
    if (!col) return -1;
    if (mode == LFO_MODE_SHW) {
        int j;
        row = (int *)malloc(sizeof(int) * ((size_t)tlen + 1));
        if (!row) { free(col); return -1; }
        lev_column(q, qlen, t, tlen, 0, col, row);
        /* first target prefix (the empty one, position -1, included) reaching the minimum */
        ed = row[0]; end = -1;
        for (j = 1; j <= tlen; j++)
            if (row[j] < ed) { ed = row[j]; end = j - 1; }
        free(row);
    } else {
        lev_column(q, qlen, t, tlen, 0, col, NULL);
        ed = col[qlen];
        end = tlen - 1; /* edlib.cpp:157-161 */
    }
    free(col);
    out->edit_distance = ed;
    out->end_location = end;
    out->n_ops = 0;
    if (want_path) {
        int rc;
        o.p = ops; o.n = 0;
        rc = path_rec(q, qlen, t, end + 1, ed, &o);
        if (rc) return rc;
        out->n_ops = o.n;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* ksw_extend2 (lib/bwa/ksw.c:380-479)                                                        */
/* ------------------------------------------------------------------------------------------ */
int lfo_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m,
                    const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w,
                    int end_bonus, int zdrop, int h0, int *qle, int *tle)
{
    /* Hd[j] feeds the diagonal: it holds H(i-1, j-1) when row i reaches column j.  Cells outside
     * the previous row's [beg,end] keep whatever an older row (or the first-row fill) left
     * there -- the reference behaves the same way and the emulation has to keep it. */
    int32_t *Hd = (int32_t *)calloc((size_t)qlen + 1, sizeof(int32_t));
    int32_t *E = (int32_t *)calloc((size_t)qlen + 1, sizeof(int32_t));
    int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int best, best_i = -1, best_j = -1, beg = 0, end = qlen, maxsc = 0, lim;
    if (!Hd || !E || h0 <= 0) { free(Hd); free(E); return -1; }
    Hd[0] = h0;
    if (qlen >= 1) Hd[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && Hd[j - 1] > e_ins; j++) Hd[j] = Hd[j - 1] - e_ins;
    for (k = 0; k < m * m; k++) if (mat[k] > maxsc) maxsc = mat[k];
    lim = (int)((double)(qlen * maxsc + end_bonus - o_ins) / e_ins + 1.);
    if (lim < 1) lim = 1;
    if (w > lim) w = lim;
    lim = (int)((double)(qlen * maxsc + end_bonus - o_del) / e_del + 1.);
    if (lim < 1) lim = 1;
    if (w > lim) w = lim;
    best = h0;
    for (i = 0; i < tlen; i++) {
        int f = 0, left, rowmax = 0, rowmax_j = -1;
        const int8_t *srow = mat + (size_t)target[i] * m;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) { left = h0 - (o_del + e_del * (i + 1)); if (left < 0) left = 0; }
        else left = 0;
        for (j = beg; j < end; j++) {
            int M = Hd[j], e = E[j], h, tt;
            Hd[j] = left;
            M = M ? M + srow[query[j]] : 0;
            h = M > e ? M : e;
            h = h > f ? h : f;
            left = h;
            if (!(rowmax > h)) rowmax_j = j; /* ties go to the later column */
            if (h > rowmax) rowmax = h;
            tt = M - oe_del; if (tt < 0) tt = 0;
            e -= e_del; if (e < tt) e = tt;
            E[j] = e;
            tt = M - oe_ins; if (tt < 0) tt = 0;
            f -= e_ins; if (f < tt) f = tt;
        }
        Hd[end] = left; E[end] = 0;
        if (rowmax == 0) break;
        if (rowmax > best) { best = rowmax; best_i = i; best_j = rowmax_j; }
        else if (zdrop > 0) {
            int di = i - best_i, dj = rowmax_j - best_j;
            if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
            else { if (best - rowmax - (dj - di) * e_ins > zdrop) break; }
        }
        for (j = beg; j < end && Hd[j] == 0 && E[j] == 0; j++) {}
        beg = j;
        for (j = end; j >= beg && Hd[j] == 0 && E[j] == 0; j--) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    free(Hd); free(E);
    if (qle) *qle = best_j + 1;
    if (tle) *tle = best_i + 1;
    return best;
}

/* ------------------------------------------------------------------------------------------ */
/* sequence helpers                                                                           */
/* ------------------------------------------------------------------------------------------ */
int lfo_pac_get(const uint8_t *pac, uint32_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

void lfo_pac2char(const uint8_t *pac, uint32_t beg, uint32_t len, char *dst)
{
    uint32_t k;
    for (k = 0; k < len; k++) dst[k] = "ACGT"[lfo_pac_get(pac, beg + k)];
}

void lfo_pac2int(const uint8_t *pac, uint32_t beg, uint32_t len, uint8_t *dst)
{
    uint32_t k;
    for (k = 0; k < len; k++) dst[k] = (uint8_t)lfo_pac_get(pac, beg + k);
}

void lfo_pack_ref(const char *seq, int64_t len, uint8_t *pac)
{
    int64_t l;
    memset(pac, 0, (size_t)(len / 4 + 1));
    for (l = 0; l < len; l++) {
        int c = seq[l] == 'A' ? 0 : seq[l] == 'C' ? 1 : seq[l] == 'G' ? 2 : 3;
        pac[l >> 2] |= (uint8_t)(c << ((~l & 3) << 1));
    }
}

static char comp_char(char c)
{ /* src/Common.cpp:34-43: case is kept, everything else becomes 'N' */
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return 'N';
    }
}

static char comp_char_upper(char c)
{ /* the table inside edlibMD_pushfront (src/LordFAST.cpp:1677-1686) folds case to upper */
    switch (c) {
    case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C'; case 'T': case 't': return 'A';
    default: return 'N';
    }
}

void lfo_revcomp(const char *src, char *dst, int len)
{
    int i;
    for (i = 0; i < len; i++) dst[i] = comp_char(src[len - 1 - i]);
    dst[len] = 0;
}

static void char2int(uint8_t *dst, const char *src, int len)
{ /* src/LordFAST.cpp:158-164, 1191-1195 */
    int i;
    for (i = 0; i < len; i++) {
        switch (src[i]) {
        case 'A': case 'a': dst[i] = 0; break; case 'C': case 'c': dst[i] = 1; break;
        case 'G': case 'g': dst[i] = 2; break; case 'T': case 't': dst[i] = 3; break;
        default: dst[i] = 4;
        }
    }
}

static void revcomp_int(uint8_t *dst, const uint8_t *src, int len)
{ /* src/LordFAST.cpp:1197-1201 (3-4 wraps to 255 exactly as there; inputs here are ACGT) */
    int i;
    for (i = 0; i < len; i++) dst[i] = (uint8_t)(3 - src[len - 1 - i]);
}

/* ------------------------------------------------------------------------------------------ */
/* alignChain_edlib (src/LordFAST.cpp:1765-2258)                                              */
/* ------------------------------------------------------------------------------------------ */

/* double-ended char buffer standing in for std::deque<char> */
typedef struct {
    char *buf;
    size_t cap, head, tail; /* live range [head, tail) */
} dq;

static int dq_init(dq *d, size_t cap) { d->buf = (char *)malloc(cap); d->cap = cap; d->head = d->tail = cap / 2; return d->buf ? 0 : -1; }
static void dq_clear(dq *d) { d->head = d->tail = d->cap / 2; }
static size_t dq_size(const dq *d) { return d->tail - d->head; }
static void dq_grow(dq *d, size_t need_front, size_t need_back)
{
    if (d->head >= need_front && d->cap - d->tail >= need_back) return;
    {
        size_t n = dq_size(d), ncap = (n + need_front + need_back) * 3 + 64;
        char *nb = (char *)malloc(ncap);
        size_t nh = need_front + (ncap - n - need_front - need_back) / 2;
        memcpy(nb + nh, d->buf + d->head, n);
        free(d->buf);
        d->buf = nb; d->cap = ncap; d->head = nh; d->tail = nh + n;
    }
}
static void dq_push_back(dq *d, char c) { dq_grow(d, 0, 1); d->buf[d->tail++] = c; }
static void dq_push_front(dq *d, char c) { dq_grow(d, 1, 0); d->buf[--d->head] = c; }
static void dq_fill_back(dq *d, size_t n, char c) { dq_grow(d, 0, n); memset(d->buf + d->tail, c, n); d->tail += n; }
static void dq_fill_front(dq *d, size_t n, char c) { dq_grow(d, n, 0); d->head -= n; memset(d->buf + d->head, c, n); }

static void cigar_push(dq *cig, const uint8_t *ops, int n, int front)
{ /* :1570-1594 */
    static const char a2c[4] = { 'M', 'I', 'D', 'M' };
    int i;
    for (i = 0; i < n; i++) { if (front) dq_push_front(cig, a2c[ops[i]]); else dq_push_back(cig, a2c[ops[i]]); }
}

static void md_push(dq *md, const char *target, const uint8_t *ops, int n, int front)
{ /* :1628-1715; the front variant complements because its target is a reverse complement */
    int i, ti = 0;
    for (i = 0; i < n; i++) {
        char c;
        switch (ops[i]) {
        case 0: c = '='; ti++; break;
        case 1: c = '-'; break;
        default: c = front ? comp_char_upper(target[ti]) : target[ti]; ti++; break;
        }
        if (front) dq_push_front(md, c); else dq_push_back(md, c);
    }
}

typedef struct { char *s; size_t n, cap; } sbuf;
static void sb_putc(sbuf *b, char c)
{
    if (b->n + 2 > b->cap) { b->cap = b->cap * 2 + 64; b->s = (char *)realloc(b->s, b->cap); }
    b->s[b->n++] = c; b->s[b->n] = 0;
}
static void sb_putnum(sbuf *b, long v) { char t[32]; int k, n = snprintf(t, sizeof t, "%ld", v); for (k = 0; k < n; k++) sb_putc(b, t[k]); }

static char *cigar_string(const dq *cig)
{ /* :1596-1626 -- leading and trailing insert runs are printed as soft clips */
    sbuf b = { NULL, 0, 0 };
    char ch = 0; long num = 0; int nops = 0; size_t i;
    sb_putc(&b, 0); b.n = 0;
    for (i = cig->head; i < cig->tail; i++) {
        if (cig->buf[i] != ch) {
            if (ch != 0) { sb_putnum(&b, num); sb_putc(&b, (nops == 0 && ch == 'I') ? 'S' : ch); nops++; }
            num = 1; ch = cig->buf[i];
        } else num++;
    }
    if (num) { sb_putnum(&b, num); sb_putc(&b, ch == 'I' ? 'S' : ch); }
    return b.s;
}

static char *md_string(const dq *md, const dq *cig)
{ /* :1717-1763 */
    sbuf b = { NULL, 0, 0 };
    long num = 0; char last = '='; size_t i, n = dq_size(md);
    sb_putc(&b, 0); b.n = 0;
    for (i = 0; i < n; i++) {
        char m = md->buf[md->head + i], c = cig->buf[cig->head + i];
        if (m == '=') { num++; last = '='; }
        else if (m == '-') { last = 'I'; }
        else if (c == 'M') { sb_putnum(&b, num); num = 0; sb_putc(&b, m); last = 'X'; }
        else if (c == 'D') {
            if (last != 'D') { sb_putnum(&b, num); num = 0; sb_putc(&b, '^'); }
            sb_putc(&b, m); last = 'D';
        }
    }
    sb_putnum(&b, num);
    return b.s;
}

static int pos2rid(const lfo_ref *r, int64_t pos)
{ /* lib/bwa/bntseq.c:349-363 */
    int left = 0, mid = 0, right = r->n_contigs;
    if (pos >= r->l_pac) return -1;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos >= r->contig_off[mid]) {
            if (mid == r->n_contigs - 1) break;
            if (pos < r->contig_off[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

typedef struct {
    const lfo_ref *ref;
    lfo_stats *st;
    uint8_t *ops; /* scratch for one alignment */
    lfo_align_out res;
} actx;

static int do_align(actx *c, const char *q, int ql, const char *t, int tl, int mode)
{
    if (c->st) { c->st->n_align++; c->st->cells_align += (int64_t)ql * tl; }
    return lfo_align(q, ql, t, tl, mode, 1, &c->res, c->ops);
}

static const int8_t CLIP_MAT[25] = { /* _pf_kswMatrix_clip, src/LordFAST.cpp:82-83, 178-187 */
    2, -16, -16, -16, 0,  -16, 2, -16, -16, 0,  -16, -16, 2, -16, 0,  -16, -16, -16, 2, 0,  0, 0, 0, 0, 0 };

static int do_extend(actx *c, int ql, const uint8_t *q, int tl, const uint8_t *t, int o_del,
                     int e_del, int o_ins, int e_ins, int w, int zdrop, int *qle, int *tle)
{
    if (c->st) { c->st->n_extend++; c->st->cells_extend += (int64_t)ql * tl; }
    return lfo_ksw_extend2(ql, q, tl, t, 5, CLIP_MAT, o_del, e_del, o_ins, e_ins, w, 0, zdrop, ql, qle, tle);
}

static void push_sam(lfo_sam *sam, int cap, int *n, const lfo_sam *tmp, const dq *cig, const dq *md, int32_t nm)
{
    if (*n >= cap) return;
    sam[*n] = *tmp;
    sam[*n].cigar = cigar_string(cig);
    sam[*n].md = md_string(md, cig);
    sam[*n].nmCount = nm;
    (*n)++;
}

int lfo_align_chain(const lfo_ref *ref, const lfo_seed *seeds, int n_seeds, const char *query,
                    int32_t read_len, int is_rev, lfo_sam *sam, int sam_cap, int *n_sam,
                    lfo_stats *stats)
{
    const int clipLen = 500, splitLen = 80;          /* :88-92 */
    const double clipSim = 0.75, splitSim = 0.40, reverseSim = 0.60;
    actx c;
    dq cig, md;
    lfo_sam tmp;
    int32_t editScore = 0, qlen, tlen;
    uint32_t chrBeg, chrEnd, qs, ts, qe, te;
    int i, rid, numAnchorsSoFar = 1, qle = 0, tle = 0;
    size_t span = (size_t)read_len + 64;
    char *qbuf, *qrev, *tbuf, *trev;
    uint8_t *qi, *ti, *qir, *tir;
    int64_t mid;

    if (n_seeds < 2) return -1;
    for (i = 0; i + 1 < n_seeds; i++) {
        size_t g = (size_t)(seeds[i + 1].tPos - seeds[i].tPos) + 64;
        if (g > span) span = g;
    }
    qbuf = (char *)malloc(span + 1); qrev = (char *)malloc(span + 1);
    tbuf = (char *)malloc(span + 1); trev = (char *)malloc(span + 1);
    qi = (uint8_t *)malloc(span); ti = (uint8_t *)malloc(span);
    qir = (uint8_t *)malloc(span); tir = (uint8_t *)malloc(span);
    c.ref = ref; c.st = stats; c.ops = (uint8_t *)malloc(2 * span);
    dq_init(&cig, 4 * span); dq_init(&md, 4 * span);
    memset(&tmp, 0, sizeof tmp);

    mid = ((int64_t)seeds[0].tPos + (int64_t)seeds[n_seeds - 1].tPos) >> 1; /* BWT.cpp:653-660 */
    rid = pos2rid(ref, mid);
    chrBeg = (uint32_t)ref->contig_off[rid];
    chrEnd = (uint32_t)(ref->contig_off[rid] + ref->contig_len[rid] - 1);

    tmp.flag = is_rev ? 16 : 0;
    tmp.pos = seeds[0].tPos;
    tmp.qStart = seeds[0].qPos;

    /* ---- head: prefix-mode alignment of the reversed read head (:1820-1899) ---- */
    qlen = (int32_t)seeds[0].qPos;
    tlen = qlen + 20;
    if (qlen > 0) {
        if ((int64_t)seeds[0].tPos - tlen >= (int64_t)chrBeg) {
            int use_first = 1;
            lfo_revcomp(query, qbuf, qlen);
            ts = seeds[0].tPos - (uint32_t)tlen;
            lfo_pac2char(ref->pac, ts, (uint32_t)tlen, tbuf);
            lfo_revcomp(tbuf, trev, tlen);
            do_align(&c, qbuf, qlen, trev, tlen, LFO_MODE_SHW);
            if (qlen > clipLen && (1 - ((float)c.res.edit_distance / qlen)) < clipSim) {
                char2int(qir, query, qlen); revcomp_int(qi, qir, qlen);
                lfo_pac2int(ref->pac, ts, (uint32_t)tlen, tir); revcomp_int(ti, tir, tlen);
                do_extend(&c, qlen, qi, tlen, ti, 0, 1, 0, 1, 40, 40, &qle, &tle);
                if (qle > 0 && qle < qlen) {
                    use_first = 0;
                    do_align(&c, qbuf, qle, trev, tle, LFO_MODE_NW);
                    cigar_push(&cig, c.ops, c.res.n_ops, 1);
                    md_push(&md, trev, c.ops, c.res.n_ops, 1);
                    editScore -= c.res.edit_distance;
                    tmp.pos = seeds[0].tPos - (uint32_t)c.res.end_location - 1;
                    tmp.qStart = seeds[0].qPos - (uint32_t)qle;
                    dq_fill_front(&cig, (size_t)(qlen - qle), 'I');
                    dq_fill_front(&md, (size_t)(qlen - qle), '-');
                }
            }
            if (use_first) {
                editScore -= c.res.edit_distance;
                cigar_push(&cig, c.ops, c.res.n_ops, 1);
                md_push(&md, trev, c.ops, c.res.n_ops, 1);
                tmp.pos = seeds[0].tPos - (uint32_t)c.res.end_location - 1;
                tmp.qStart = 0;
            }
        } else {
            dq_fill_front(&cig, (size_t)qlen, 'I');
            dq_fill_front(&md, (size_t)qlen, '-');
        }
    }

    /* ---- anchors and the gaps between them (:1901-2137) ---- */
    for (i = 0; i < n_seeds - 1; i++) {
        dq_fill_back(&cig, seeds[i].len, 'M');
        dq_fill_back(&md, seeds[i].len, '=');
        qs = seeds[i].qPos + seeds[i].len; ts = seeds[i].tPos + seeds[i].len;
        qe = seeds[i + 1].qPos; te = seeds[i + 1].tPos;
        qlen = (int32_t)(qe - qs); tlen = (int32_t)(te - ts);
        if (qlen > 0 && tlen > 0) {
            int plain = 1;
            lfo_pac2char(ref->pac, ts, (uint32_t)tlen, tbuf);
            do_align(&c, query + qs, qlen, tbuf, tlen, LFO_MODE_NW);
            if (abs(qlen - tlen) >= splitLen && (1 - ((float)c.res.edit_distance / qlen)) < splitSim) {
                uint32_t qs2, ts2, qe2, te2; int32_t ql2, tl2;
                char2int(qi, query + qs, qlen);
                lfo_pac2int(ref->pac, ts, (uint32_t)tlen, ti);
                do_extend(&c, qlen, qi, tlen, ti, 8, 1, 4, 1, 100, 200, &qle, &tle);
                qs2 = qs + (uint32_t)qle; ts2 = ts + (uint32_t)tle;
                char2int(qir, query + qs, qlen); revcomp_int(qi, qir, qlen);
                lfo_pac2int(ref->pac, ts, (uint32_t)tlen, tir); revcomp_int(ti, tir, tlen);
                do_extend(&c, qlen, qi, tlen, ti, 8, 1, 4, 1, 100, 200, &qle, &tle);
                qe2 = qe - (uint32_t)qle; te2 = te - (uint32_t)tle;
                tl2 = (int32_t)(te2 - ts2); ql2 = (int32_t)(qe2 - qs2);
                if (qs2 < qe2 || ts2 < te2) { /* extensions do not cross: split here (:1995) */
                    plain = 0;
                    if (qs2 > qs || ts2 > ts) {
                        do_align(&c, query + qs, (int)(qs2 - qs), tbuf, (int)(ts2 - ts), LFO_MODE_NW);
                        cigar_push(&cig, c.ops, c.res.n_ops, 0);
                        md_push(&md, tbuf, c.ops, c.res.n_ops, 0);
                        editScore -= c.res.edit_distance;
                    }
                    dq_fill_back(&cig, (size_t)((uint32_t)read_len - qs2), 'I');
                    dq_fill_back(&md, (size_t)((uint32_t)read_len - qs2), '-');
                    tmp.posEnd = ts2; tmp.qEnd = qs2;
                    if (numAnchorsSoFar > 1) push_sam(sam, sam_cap, n_sam, &tmp, &cig, &md, editScore);
                    dq_clear(&cig); dq_clear(&md); editScore = 0;
                    if (qs2 < qe2 && ts2 < te2) { /* middle part: inversion test (:2034-2077) */
                        lfo_align_out fwd;
                        lfo_pac2char(ref->pac, ts2, (uint32_t)tl2, tbuf);
                        do_align(&c, query + qs2, ql2, tbuf, tl2, LFO_MODE_NW);
                        fwd = c.res;
                        lfo_revcomp(query + qs2, qrev, ql2);
                        do_align(&c, qrev, ql2, tbuf, tl2, LFO_MODE_NW);
                        if ((1 - ((double)c.res.edit_distance / ql2)) > (1 - ((double)fwd.edit_distance / ql2))
                            && (1 - ((double)c.res.edit_distance / ql2)) > reverseSim) {
                            tmp.flag = is_rev ? 0 : 16;
                            tmp.pos = ts2; tmp.qStart = qs2; tmp.posEnd = te2; tmp.qEnd = qe2;
                            dq_fill_back(&cig, qs2, 'I');
                            dq_fill_back(&md, qs2, '-');
                            cigar_push(&cig, c.ops, c.res.n_ops, 0);
                            md_push(&md, tbuf, c.ops, c.res.n_ops, 0);
                            editScore -= c.res.edit_distance;
                            dq_fill_back(&cig, (size_t)((uint32_t)read_len - qe2), 'I');
                            dq_fill_front(&md, (size_t)((uint32_t)read_len - qe2), '-'); /* sic, :2057 */
                            push_sam(sam, sam_cap, n_sam, &tmp, &cig, &md, editScore);
                            dq_clear(&cig); dq_clear(&md); editScore = 0;
                        }
                    }
                    if (qe2 < qe || te2 < te) { /* second part, aligned right-to-left (:2080-2090) */
                        lfo_pac2char(ref->pac, ts, (uint32_t)tlen, tbuf); /* tbuf may hold the middle part */
                        lfo_revcomp(query + qs, qbuf, qlen);
                        lfo_revcomp(tbuf, trev, tlen);
                        do_align(&c, qbuf, (int)(qe - qe2), trev, (int)(te - te2), LFO_MODE_NW);
                        cigar_push(&cig, c.ops, c.res.n_ops, 1);
                        md_push(&md, trev, c.ops, c.res.n_ops, 1);
                        editScore -= c.res.edit_distance;
                    }
                    dq_fill_front(&cig, qe2, 'I');
                    dq_fill_front(&md, qe2, '-');
                    tmp.flag = is_rev ? 16 : 0;
                    tmp.pos = te2; tmp.qStart = qe2;
                    numAnchorsSoFar = 0;
                }
            }
            if (plain) {
                editScore -= c.res.edit_distance;
                cigar_push(&cig, c.ops, c.res.n_ops, 0);
                md_push(&md, tbuf, c.ops, c.res.n_ops, 0);
            }
        } else if (qlen > 0) {
            dq_fill_back(&cig, (size_t)qlen, 'I');
            dq_fill_back(&md, (size_t)qlen, '-');
            editScore -= qlen;
        } else {
            int j;
            dq_fill_back(&cig, (size_t)tlen, 'D');
            lfo_pac2char(ref->pac, ts, (uint32_t)tlen, tbuf);
            for (j = 0; j < tlen; j++) dq_push_back(&md, tbuf[j]);
            editScore -= tlen;
        }
        numAnchorsSoFar++;
    }
    dq_fill_back(&cig, seeds[i].len, 'M');
    dq_fill_back(&md, seeds[i].len, '=');
    tmp.posEnd = seeds[i].tPos + seeds[i].len - 1;
    tmp.qEnd = seeds[i].qPos + seeds[i].len - 1;

    /* ---- tail: prefix-mode alignment of the rest of the read (:2157-2230) ---- */
    qs = seeds[i].qPos + seeds[i].len;
    qlen = read_len - (int32_t)qs;
    tlen = qlen + 20;
    if (qlen > 0) {
        if (seeds[i].tPos + seeds[i].len + (uint32_t)tlen - 1 <= chrEnd) {
            int use_first = 1;
            ts = seeds[i].tPos + seeds[i].len;
            lfo_pac2char(ref->pac, ts, (uint32_t)tlen, tbuf);
            do_align(&c, query + qs, qlen, tbuf, tlen, LFO_MODE_SHW);
            if (qlen > clipLen && (1 - ((float)c.res.edit_distance / qlen)) < clipSim) {
                char2int(qi, query + qs, qlen);
                lfo_pac2int(ref->pac, ts, (uint32_t)tlen, ti);
                do_extend(&c, qlen, qi, tlen, ti, 0, 1, 0, 1, 40, 40, &qle, &tle);
                if (qle > 0 && qle < qlen) {
                    use_first = 0;
                    do_align(&c, query + qs, qle, tbuf, tle, LFO_MODE_NW);
                    cigar_push(&cig, c.ops, c.res.n_ops, 0);
                    md_push(&md, tbuf, c.ops, c.res.n_ops, 0);
                    editScore -= c.res.edit_distance;
                    tmp.posEnd = ts + (uint32_t)c.res.end_location;
                    tmp.qEnd = qs + (uint32_t)qle;
                    dq_fill_back(&cig, (size_t)(qlen - qle), 'I');
                    dq_fill_back(&md, (size_t)(qlen - qle), '-');
                }
            }
            if (use_first) {
                editScore -= c.res.edit_distance;
                cigar_push(&cig, c.ops, c.res.n_ops, 0);
                md_push(&md, tbuf, c.ops, c.res.n_ops, 0);
                tmp.posEnd = ts + (uint32_t)c.res.end_location;
                tmp.qEnd = (uint32_t)read_len;
            }
        } else {
            dq_fill_back(&cig, (size_t)qlen, 'I');
            dq_fill_back(&md, (size_t)qlen, '-');
        }
    }
    push_sam(sam, sam_cap, n_sam, &tmp, &cig, &md, editScore);

    free(qbuf); free(qrev); free(tbuf); free(trev); free(qi); free(ti); free(qir); free(tir);
    free(c.ops); free(cig.buf); free(md.buf);
    return 0;
}

void lfo_free_sam(lfo_sam *sam, int n)
{
    int i;
    for (i = 0; i < n; i++) { free(sam[i].cigar); free(sam[i].md); sam[i].cigar = sam[i].md = NULL; }
}
