// blacs.cpp -- BLACS setup API kept for the Fortran caller: process grid
// contexts with the reference's row-major / column-major rank maps (BLACS/SRC/blacs_init_.c:23-40,
// blacs_map_.c:84-141) and the three scopes (row, column, all).  Data-path broadcasts/combines of the
// LU path do NOT go through here; they are NCCL (lu_dist.cu).
#include "common.h"
#include "ncclw.h"

#include <mutex>

namespace slb {

static std::vector<Grid *> g_grids;     // index = context handle (BI_MyContxts analogue)
static std::mutex g_gmu;

Grid *grid_of(int ictxt)
{
    std::lock_guard<std::mutex> lk(g_gmu);
    if (ictxt < 0 || ictxt >= (int)g_grids.size()) return nullptr;
    Grid *g = g_grids[ictxt];
    return (g && g->valid) ? g : nullptr;
}

int grid_scope_size(Grid *g, char scope)
{ return scope == 'R' ? g->npcol : (scope == 'C' ? g->nprow : g->nprow * g->npcol); }
int grid_scope_index(Grid *g, char scope)
{ return scope == 'R' ? g->mycol : (scope == 'C' ? g->myrow : g->myrow * g->npcol + g->mycol); }

void grid_allgather(Grid *g, char scope, const void *in, void *out, size_t len)
{
    int n = grid_scope_size(g, scope), idx = grid_scope_index(g, scope);
    uint64_t seq, coord, sc;
    if (scope == 'R') { seq = g->seq_row++; coord = (uint64_t)g->myrow; sc = 1; }
    else if (scope == 'C') { seq = g->seq_col++; coord = (uint64_t)g->mycol; sc = 2; }
    else { seq = g->seq_all++; coord = 0; sc = 0; }
    uint64_t key = ((uint64_t)(g->uid + 1) << 50) | (sc << 48) | (coord << 34) | (seq & 0x3ffffffffULL);
    hc_allgather(key, n, idx, in, out, len);
}
void grid_barrier(Grid *g, char scope)
{
    int n = grid_scope_size(g, scope);
    if (n <= 1) return;
    std::vector<char> tmp((size_t)n);
    char z = 0; grid_allgather(g, scope, &z, tmp.data(), 1);
}
int grid_imin(Grid *g, char scope, int v)
{
    int n = grid_scope_size(g, scope);
    if (n <= 1) return v;
    std::vector<int> all((size_t)n); grid_allgather(g, scope, &v, all.data(), sizeof(int));
    for (int x : all) if (x < v) v = x;
    return v;
}
int grid_imax(Grid *g, char scope, int v)
{
    int n = grid_scope_size(g, scope);
    if (n <= 1) return v;
    std::vector<int> all((size_t)n); grid_allgather(g, scope, &v, all.data(), sizeof(int));
    for (int x : all) if (x > v) v = x;
    return v;
}

static char scope_char(const char *scope)
{
    char c = scope ? (char)(scope[0] & ~0x20) : 'A';
    return (c == 'R' || c == 'C') ? c : 'A';
}

static int new_grid(const std::vector<int> &pmap, int nprow, int npcol)
{
    static int uid_counter = 0;
    int me = hc_rank();
    Grid *g = new Grid();
    g->uid = ++uid_counter;
    g->valid = true; g->nprow = nprow; g->npcol = npcol; g->pmap = pmap;
    for (int r = 0; r < nprow; ++r) for (int c = 0; c < npcol; ++c)
        if (pmap[(size_t)r * npcol + c] == me) { g->myrow = r; g->mycol = c; }
    std::lock_guard<std::mutex> lk(g_gmu);
    int slot = -1;
    for (size_t i = 0; i < g_grids.size(); ++i) if (!g_grids[i]) { slot = (int)i; break; }   // first free slot
    if (slot < 0) { g_grids.push_back(nullptr); slot = (int)g_grids.size() - 1; }
    g->ctxt = slot;
    g_grids[slot] = g;
    return slot;
}

}  // namespace slb

using namespace slb;

extern "C" {

// BLACS/SRC/blacs_pinfo_.c:14-27
void blacs_pinfo_(int *mypnum, int *nprocs) { *mypnum = hc_rank(); *nprocs = hc_size(); }

// BLACS/SRC/blacs_get_.c: what=0 -> default system context handle; what=10 -> system context of a grid.
void blacs_get_(const int *ictxt, const int *what, int *val)
{
    (void)ictxt;
    switch (*what) {
    case 0: *val = 0; break;
    case 10: *val = 0; break;
    default: *val = 0; break;
    }
}
void blacs_set_(const int *ictxt, const int *what, const int *val) { (void)ictxt; (void)what; (void)val; }

// BLACS/SRC/blacs_map_.c:84-141.  usermap(ldumap, npcol) column-major: usermap[r + c*ldumap] = world rank.
void blacs_gridmap_(int *ictxt, const int *usermap, const int *ldumap, const int *nprow, const int *npcol)
{
    int P = *nprow, Q = *npcol, np = hc_size();
    if (P < 1 || Q < 1) fatal("BLACS_GRIDMAP: illegal grid (%d x %d)", P, Q);
    if (P * Q > np) fatal("BLACS_GRIDMAP: grid too big: %d x %d > %d processes", P, Q, np);
    std::vector<int> pmap((size_t)P * Q);
    for (int r = 0; r < P; ++r) for (int c = 0; c < Q; ++c) pmap[(size_t)r * Q + c] = usermap[r + (size_t)c * *ldumap];
    int slot = new_grid(pmap, P, Q);
    Grid *g = grid_of(slot);
    *ictxt = g->myrow >= 0 ? slot : -1;          // NOTINCONTEXT (blacs_map_.c:72-77)
    if (g->myrow < 0) { std::lock_guard<std::mutex> lk(g_gmu); delete g_grids[slot]; g_grids[slot] = nullptr; }
}

// BLACS/SRC/blacs_init_.c:23-40: 'C'/'c' => column-major rank map, anything else row-major.
void blacs_gridinit_(int *ictxt, const char *order, const int *nprow, const int *npcol)
{
    int P = *nprow, Q = *npcol;
    if (P < 1 || Q < 1) fatal("BLACS_GRIDINIT: illegal grid (%d x %d)", P, Q);
    std::vector<int> umap((size_t)P * Q);
    bool colmajor = order && (order[0] == 'C' || order[0] == 'c');
    for (int r = 0; r < P; ++r) for (int c = 0; c < Q; ++c)
        umap[r + (size_t)c * P] = colmajor ? (c * P + r) : (r * Q + c);
    blacs_gridmap_(ictxt, umap.data(), &P, nprow, npcol);
}

// BLACS/SRC/blacs_info_.c
void blacs_gridinfo_(const int *ictxt, int *nprow, int *npcol, int *myrow, int *mycol)
{
    Grid *g = grid_of(*ictxt);
    if (!g || g->myrow < 0) { *nprow = *npcol = *myrow = *mycol = -1; return; }
    *nprow = g->nprow; *npcol = g->npcol; *myrow = g->myrow; *mycol = g->mycol;
}

// BLACS/SRC/blacs_grid_.c
void blacs_gridexit_(const int *ictxt)
{
    Grid *g = grid_of(*ictxt);
    if (!g) { fprintf(stderr, "BLACS_GRIDEXIT: trying to exit non-existent context %d\n", *ictxt); return; }
    if (g->nccl) { nccl_destroy(g->nccl); g->nccl = nullptr; }
    std::lock_guard<std::mutex> lk(g_gmu);
    g_grids[*ictxt] = nullptr;
    delete g;
}

void blacs_exit_(const int *notdone)
{
    {
        std::vector<int> live;
        { std::lock_guard<std::mutex> lk(g_gmu); for (size_t i = 0; i < g_grids.size(); ++i) if (g_grids[i]) live.push_back((int)i); }
        for (int c : live) blacs_gridexit_(&c);
    }
    if (*notdone == 0) hc_shutdown();
}

void blacs_abort_(const int *ictxt, const int *errnum)
{
    int P, Q, r, c; blacs_gridinfo_(ictxt, &P, &Q, &r, &c);
    fprintf(stderr, "{%d,%d}, pnum=%d, Contxt=%d, killed other procs, exiting with error #%d.\n\n", r, c, hc_rank(), *ictxt, *errnum);
    fflush(stderr);
    _Exit(*errnum ? *errnum : 1);
}

// BLACS/SRC/blacs_barr_.c:16-26
void blacs_barrier_(const int *ictxt, const char *scope)
{
    Grid *g = grid_of(*ictxt);
    if (!g || !g->in_grid()) return;
    grid_barrier(g, scope_char(scope));
}

int blacs_pnum_(const int *ictxt, const int *prow, const int *pcol)
{
    Grid *g = grid_of(*ictxt);
    if (!g || *prow < 0 || *prow >= g->nprow || *pcol < 0 || *pcol >= g->npcol) return -1;
    return g->pmap[(size_t)*prow * g->npcol + *pcol];
}
void blacs_pcoord_(const int *ictxt, const int *pnum, int *prow, int *pcol)
{
    Grid *g = grid_of(*ictxt);
    *prow = *pcol = -1;
    if (!g) return;
    for (int r = 0; r < g->nprow; ++r) for (int c = 0; c < g->npcol; ++c)
        if (g->pmap[(size_t)r * g->npcol + c] == *pnum) { *prow = r; *pcol = c; }
}

// BLACS/SRC/igamn2d_.c / igamx2d_.c: element-wise min / max of an m x n int matrix over a scope; result on
// (rdest,cdest) or everywhere when rdest == -1.  RA/CA coordinate outputs are not produced (callers on the
// LU path pass rcflag = -1, SRC/pdgetrf.f:299).
static void igam(const int *ictxt, const char *scope, const int *m, const int *n, int *a, const int *lda, bool want_min)
{
    Grid *g = grid_of(*ictxt);
    if (!g || !g->in_grid()) return;
    char sc = scope_char(scope);
    int np = grid_scope_size(g, sc);
    if (np <= 1) return;
    size_t cnt = (size_t)*m * *n;
    std::vector<int> mine(cnt), all(cnt * np);
    for (int j = 0; j < *n; ++j) for (int i = 0; i < *m; ++i) mine[i + (size_t)j * *m] = a[i + (size_t)j * *lda];
    grid_allgather(g, sc, mine.data(), all.data(), cnt * sizeof(int));
    for (size_t e = 0; e < cnt; ++e) {
        int v = all[e];
        for (int p = 1; p < np; ++p) { int x = all[(size_t)p * cnt + e]; v = want_min ? (x < v ? x : v) : (x > v ? x : v); }
        mine[e] = v;
    }
    for (int j = 0; j < *n; ++j) for (int i = 0; i < *m; ++i) a[i + (size_t)j * *lda] = mine[i + (size_t)j * *m];
}
void igamn2d_(const int *ictxt, const char *scope, const char *top, const int *m, const int *n, int *a, const int *lda,
              int *ra, int *ca, const int *rcflag, const int *rdest, const int *cdest)
{ (void)top; (void)ra; (void)ca; (void)rcflag; (void)rdest; (void)cdest; igam(ictxt, scope, m, n, a, lda, true); }
void igamx2d_(const int *ictxt, const char *scope, const char *top, const int *m, const int *n, int *a, const int *lda,
              int *ra, int *ca, const int *rcflag, const int *rdest, const int *cdest)
{ (void)top; (void)ra; (void)ca; (void)rcflag; (void)rdest; (void)cdest; igam(ictxt, scope, m, n, a, lda, false); }

// TOOLS/SL_init.f
void sl_init_(int *ictxt, const int *nprow, const int *npcol)
{
    int me, np; blacs_pinfo_(&me, &np);
    int m1 = -1, z = 0; blacs_get_(&m1, &z, ictxt);
    blacs_gridinit_(ictxt, "Row-major", nprow, npcol);
}

// C twins
void Cblacs_pinfo(int *mypnum, int *nprocs) { blacs_pinfo_(mypnum, nprocs); }
void Cblacs_get(int ictxt, int what, int *val) { blacs_get_(&ictxt, &what, val); }
void Cblacs_gridinit(int *ictxt, const char *order, int nprow, int npcol) { blacs_gridinit_(ictxt, order, &nprow, &npcol); }
void Cblacs_gridinfo(int ictxt, int *nprow, int *npcol, int *myrow, int *mycol) { blacs_gridinfo_(&ictxt, nprow, npcol, myrow, mycol); }
void Cblacs_gridexit(int ictxt) { blacs_gridexit_(&ictxt); }
void Cblacs_exit(int notdone) { blacs_exit_(&notdone); }
void Cblacs_barrier(int ictxt, const char *scope) { blacs_barrier_(&ictxt, scope); }
int  Cblacs_pnum(int ictxt, int prow, int pcol) { return blacs_pnum_(&ictxt, &prow, &pcol); }
void Cblacs_pcoord(int ictxt, int pnum, int *prow, int *pcol) { blacs_pcoord_(&ictxt, &pnum, prow, pcol); }

}  // extern "C"
