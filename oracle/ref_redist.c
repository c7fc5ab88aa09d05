/*
 * ref_redist.c -- harness that runs the REFERENCE's own PDGEMR2D (REDIST/SRC/pdgemr.c, pgemraux.c, pdgemr2.c, compiled where they
 * lie under /root/reference by oracle/Makefile into oracle/_ref/libref_redist.so) so that the redistribution tests are pinned against
 * the reference itself, not only against this repository's restatement.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference code is SPMD over the BLACS; there is no MPI here, so this file provides the dozen BLACS
 * entry points those three files use ("mini-BLACS") with one THREAD per BLACS process inside one OS process: grids are tables,
 * point-to-point messages are mailboxes (sends are buffered like the BLACS'), the one combine (IGAMN2D) is a barrier exchange.
 * Nothing of the reference is copied: its sources are compiled in place and linked with this harness.
 *
 *   ref_pdgemr2d_run(): every thread builds its block-cyclic piece of the global A, calls the reference's Cpdgemr2d and hands back
 *   its piece of B.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int desctype, ctxt, m, n, nbrow, nbcol, sprow, spcol, lda; } MDESC;       /* REDIST/SRC/pdgemr.c:160-170 */
void Cpdgemr2d(int m, int n, double *a, int ia, int ja, MDESC *ma, double *b, int ib, int jb, MDESC *mb, int gctxt);   /* the reference */
void Cpzgemr2d(int m, int n, void *a, int ia, int ja, MDESC *ma, void *b, int ib, int jb, MDESC *mb, int gctxt);       /* REDIST/SRC/pzgemr.c */

/* ---- mini-BLACS ------------------------------------------------------------------------------------------------------------- */
#define MAXCTX 16
#define MAXP 64
typedef struct { int valid, nprow, npcol, pmap[MAXP]; pthread_barrier_t bar; int *slots; } Ctx;
static Ctx g_ctx[MAXCTX];
static int g_np;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;
static __thread int t_rank, t_nctx;

typedef struct Msg { struct Msg *next; size_t len; char data[]; } Msg;
static Msg *g_box[MAXP][MAXP], *g_boxtail[MAXP][MAXP];

static int coords(const Ctx *c, int pnum, int *r, int *col)
{
    for (int i = 0; i < c->nprow * c->npcol; ++i) if (c->pmap[i] == pnum) { *r = i / c->npcol; *col = i % c->npcol; return 1; }
    *r = *col = -1; return 0;
}
void Cblacs_pinfo(int *mypnum, int *nprocs) { *mypnum = t_rank; *nprocs = g_np; }
void Cblacs_get(int ctxt, int what, int *val) { (void)ctxt; (void)what; *val = 0; }
static void new_grid(int *ctxt, const int *pmap, int nprow, int npcol)
{
    const int h = 1 + t_nctx++;                                   /* every thread creates its grids in the same order */
    pthread_mutex_lock(&g_mu);
    Ctx *c = &g_ctx[h];
    if (!c->valid) {
        c->nprow = nprow; c->npcol = npcol;
        memcpy(c->pmap, pmap, sizeof(int) * (size_t)(nprow * npcol));
        pthread_barrier_init(&c->bar, NULL, (unsigned)(nprow * npcol));
        c->valid = 1;
    }
    pthread_mutex_unlock(&g_mu);
    *ctxt = h;
}
void Cblacs_gridinit(int *ctxt, char *order, int nprow, int npcol)
{
    int pmap[MAXP];
    for (int i = 0; i < nprow * npcol; ++i) pmap[i] = (order[0] == 'C' || order[0] == 'c') ? (i % npcol) * nprow + i / npcol : i;
    new_grid(ctxt, pmap, nprow, npcol);
}
void Cblacs_gridmap(int *ctxt, int *usermap, int ldumap, int nprow, int npcol)
{
    int pmap[MAXP];
    for (int r = 0; r < nprow; ++r) for (int c = 0; c < npcol; ++c) pmap[r * npcol + c] = usermap[r + c * ldumap];
    new_grid(ctxt, pmap, nprow, npcol);
}
void Cblacs_gridinfo(int ctxt, int *nprow, int *npcol, int *myrow, int *mycol)
{
    *nprow = *npcol = *myrow = *mycol = -1;
    if (ctxt <= 0 || ctxt >= MAXCTX || !g_ctx[ctxt].valid) return;
    if (!coords(&g_ctx[ctxt], t_rank, myrow, mycol)) return;     /* not a member: all -1 (BLACS/SRC/blacs_info_.c) */
    *nprow = g_ctx[ctxt].nprow; *npcol = g_ctx[ctxt].npcol;
}
int Cblacs_pnum(int ctxt, int prow, int pcol) { return g_ctx[ctxt].pmap[prow * g_ctx[ctxt].npcol + pcol]; }
void Cblacs_pcoord(int ctxt, int pnum, int *prow, int *pcol) { coords(&g_ctx[ctxt], pnum, prow, pcol); }
void Cblacs_gridexit(int ctxt) { (void)ctxt; }
void Cblacs_exit(int notdone) { (void)notdone; }

static void send_bytes(int dst, const void *p, size_t len)
{
    Msg *m = malloc(sizeof(Msg) + len);
    m->next = NULL; m->len = len; memcpy(m->data, p, len);
    pthread_mutex_lock(&g_mu);
    if (g_boxtail[t_rank][dst]) g_boxtail[t_rank][dst]->next = m; else g_box[t_rank][dst] = m;
    g_boxtail[t_rank][dst] = m;
    pthread_cond_broadcast(&g_cv);
    pthread_mutex_unlock(&g_mu);
}
static void recv_bytes(int src, void *p, size_t len)
{
    pthread_mutex_lock(&g_mu);
    while (!g_box[src][t_rank]) pthread_cond_wait(&g_cv, &g_mu);
    Msg *m = g_box[src][t_rank];
    g_box[src][t_rank] = m->next;
    if (!m->next) g_boxtail[src][t_rank] = NULL;
    pthread_mutex_unlock(&g_mu);
    if (m->len != len) { fprintf(stderr, "mini-BLACS: message of %zu bytes received as %zu\n", m->len, len); abort(); }
    memcpy(p, m->data, len);
    free(m);
}
/* m x n general matrices, column-major with leading dimension lda, packed on the wire (BLACS/SRC/dgesd2d_.c) */
static void gesd(int ctxt, int m, int n, const char *a, int lda, int rdest, int cdest, size_t es)
{
    char *buf = malloc((size_t)m * n * es + 1);
    for (int j = 0; j < n; ++j) memcpy(buf + (size_t)j * m * es, a + (size_t)j * lda * es, (size_t)m * es);
    send_bytes(Cblacs_pnum(ctxt, rdest, cdest), buf, (size_t)m * n * es);
    free(buf);
}
static void gerv(int ctxt, int m, int n, char *a, int lda, int rsrc, int csrc, size_t es)
{
    char *buf = malloc((size_t)m * n * es + 1);
    recv_bytes(Cblacs_pnum(ctxt, rsrc, csrc), buf, (size_t)m * n * es);
    for (int j = 0; j < n; ++j) memcpy(a + (size_t)j * lda * es, buf + (size_t)j * m * es, (size_t)m * es);
    free(buf);
}
void Cdgesd2d(int ctxt, int m, int n, double *a, int lda, int rdest, int cdest) { gesd(ctxt, m, n, (const char *)a, lda, rdest, cdest, 8); }
void Cdgerv2d(int ctxt, int m, int n, double *a, int lda, int rsrc, int csrc) { gerv(ctxt, m, n, (char *)a, lda, rsrc, csrc, 8); }
void Czgesd2d(int ctxt, int m, int n, void *a, int lda, int rdest, int cdest) { gesd(ctxt, m, n, (const char *)a, lda, rdest, cdest, 16); }
void Czgerv2d(int ctxt, int m, int n, void *a, int lda, int rsrc, int csrc) { gerv(ctxt, m, n, (char *)a, lda, rsrc, csrc, 16); }
void Cigesd2d(int ctxt, int m, int n, int *a, int lda, int rdest, int cdest) { gesd(ctxt, m, n, (const char *)a, lda, rdest, cdest, 4); }
void Cigerv2d(int ctxt, int m, int n, int *a, int lda, int rsrc, int csrc) { gerv(ctxt, m, n, (char *)a, lda, rsrc, csrc, 4); }
/* element-wise minimum over ALL processes of the context, result everywhere (BLACS/SRC/igamn2d_.c with rdest = -1) */
void Cigamn2d(int ctxt, char *scope, char *top, int m, int n, int *a, int lda, int *ra, int *ca, int rcflag, int rdest, int cdest)
{
    (void)scope; (void)top; (void)rcflag; (void)rdest; (void)cdest;
    Ctx *c = &g_ctx[ctxt];
    const int np = c->nprow * c->npcol, cnt = m * n;
    int r, col; coords(c, t_rank, &r, &col);
    const int me = r * c->npcol + col;
    pthread_barrier_wait(&c->bar);
    if (me == 0) c->slots = malloc(sizeof(int) * (size_t)np * cnt);
    pthread_barrier_wait(&c->bar);
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) c->slots[(size_t)me * cnt + i + j * m] = a[i + j * lda];
    pthread_barrier_wait(&c->bar);
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) {
        int v = c->slots[i + j * m];
        for (int p = 1; p < np; ++p) if (c->slots[(size_t)p * cnt + i + j * m] < v) v = c->slots[(size_t)p * cnt + i + j * m];
        a[i + j * lda] = v;
        if (ra) ra[i + j * m] = 0;
        if (ca) ca[i + j * m] = 0;
    }
    pthread_barrier_wait(&c->bar);
    if (me == 0) { free(c->slots); c->slots = NULL; }
}

/* ---- the run ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int rank, np, m, n, ia, ja, ib, jb;
    int pa, qa, ma, na, mba, nba, rsa, csa;
    int pb, qb, mb, nb, mbb, nbb, rsb, csb;
    const double *aglob;            /* ma x na, column-major (es / 8 doubles per element) */
    double fill;                    /* what B holds outside sub(B) (every double of an element) */
    double *bout; int64_t bstride;  /* rank r returns its local B (lld = max(1, LOCr) x LOCc, packed) at bout + r * bstride (in doubles) */
    int *bdims;                     /* 2 ints per rank: LOCr, LOCc of B (0, 0 when not in B's grid) */
    int es;                         /* element size: 8 (PDGEMR2D) or 16 (PZGEMR2D) */
} Job;

static int numroc(int n, int nb, int iproc, int isrc, int nprocs)
{
    const int mydist = (nprocs + iproc - isrc) % nprocs, nblocks = n / nb;
    int v = (nblocks / nprocs) * nb;
    const int extra = nblocks % nprocs;
    if (mydist < extra) v += nb; else if (mydist == extra) v += n % nb;
    return v;
}
static void *worker(void *arg)
{
    Job *j = (Job *)arg;
    t_rank = j->rank; t_nctx = 0;
    int gctxt, ctxa, ctxb, p, q, r, c;
    Cblacs_get(0, 0, &gctxt); Cblacs_gridinit(&gctxt, "R", 1, j->np);
    Cblacs_get(0, 0, &ctxa); Cblacs_gridinit(&ctxa, "R", j->pa, j->qa);
    Cblacs_get(0, 0, &ctxb); Cblacs_gridinit(&ctxb, "R", j->pb, j->qb);
    MDESC da = { 1, -1, j->ma, j->na, j->mba, j->nba, j->rsa, j->csa, 1 }, db = { 1, -1, j->mb, j->nb, j->mbb, j->nbb, j->rsb, j->csb, 1 };
    double *al = NULL, *bl = NULL;
    const int w = j->es / 8;                                      /* doubles per element */
    Cblacs_gridinfo(ctxa, &p, &q, &r, &c);
    if (r >= 0) {
        const int ml = numroc(j->ma, j->mba, r, j->rsa, p), nl = numroc(j->na, j->nba, c, j->csa, q), lld = ml > 0 ? ml : 1;
        al = malloc(sizeof(double) * w * (size_t)lld * (size_t)(nl > 0 ? nl : 1));
        for (int gj = 0; gj < j->na; ++gj) {
            if ((j->csa + gj / j->nba) % q != c) continue;
            const int lj = (gj / (j->nba * q)) * j->nba + gj % j->nba;
            for (int gi = 0; gi < j->ma; ++gi) {
                if ((j->rsa + gi / j->mba) % p != r) continue;
                for (int d = 0; d < w; ++d) al[w * ((gi / (j->mba * p)) * j->mba + gi % j->mba + (size_t)lj * lld) + d] = j->aglob[w * (gi + (size_t)gj * j->ma) + d];
            }
        }
        da.ctxt = ctxa; da.lda = lld;
    }
    Cblacs_gridinfo(ctxb, &p, &q, &r, &c);
    int mlb = 0, nlb = 0, lldb = 1;
    if (r >= 0) {
        mlb = numroc(j->mb, j->mbb, r, j->rsb, p); nlb = numroc(j->nb, j->nbb, c, j->csb, q); lldb = mlb > 0 ? mlb : 1;
        bl = malloc(sizeof(double) * w * (size_t)lldb * (size_t)(nlb > 0 ? nlb : 1));
        for (size_t e = 0; e < (size_t)w * lldb * (size_t)(nlb > 0 ? nlb : 1); ++e) bl[e] = j->fill;
        db.ctxt = ctxb; db.lda = lldb;
    }
    double dummy[2] = { 0.0, 0.0 };
    if (j->es == 8) Cpdgemr2d(j->m, j->n, al ? al : dummy, j->ia, j->ja, &da, bl ? bl : dummy, j->ib, j->jb, &db, gctxt);
    else Cpzgemr2d(j->m, j->n, al ? (void *)al : (void *)dummy, j->ia, j->ja, &da, bl ? (void *)bl : (void *)dummy, j->ib, j->jb, &db, gctxt);
    j->bdims[2 * j->rank] = mlb; j->bdims[2 * j->rank + 1] = nlb;
    if (bl) memcpy(j->bout + (size_t)j->rank * j->bstride, bl, sizeof(double) * w * (size_t)lldb * (size_t)(nlb > 0 ? nlb : 1));
    free(al); free(bl);
    return NULL;
}

/* Runs the reference redistribution of sub(A) = A(ia:ia+m-1, ja:ja+n-1) (A: ma x na on a pa x qa grid, blocks mba x nba from process
 * (rsa, csa)) into sub(B) of a B filled with `fill` (mb x nb on a pb x qb grid, ...), with np = max(pa qa, pb qb) BLACS processes.
 * Returns 0. */
int ref_pdgemr2d_run(int np, int m, int n, int ia, int ja, int ib, int jb,
                     int pa, int qa, int ma, int na, int mba, int nba, int rsa, int csa,
                     int pb, int qb, int mb, int nb, int mbb, int nbb, int rsb, int csb,
                     const double *aglob, double fill, double *bout, int64_t bstride, int *bdims, int es)
{
    if (np > MAXP) return -1;
    memset(g_ctx, 0, sizeof(g_ctx)); memset(g_box, 0, sizeof(g_box)); memset(g_boxtail, 0, sizeof(g_boxtail));
    g_np = np;
    pthread_t th[MAXP]; Job jobs[MAXP];
    for (int r = 0; r < np; ++r) {
        jobs[r] = (Job){ r, np, m, n, ia, ja, ib, jb, pa, qa, ma, na, mba, nba, rsa, csa, pb, qb, mb, nb, mbb, nbb, rsb, csb, aglob, fill, bout, bstride, bdims, es };
        pthread_create(&th[r], NULL, worker, &jobs[r]);
    }
    for (int r = 0; r < np; ++r) pthread_join(th[r], NULL);
    return 0;
}
