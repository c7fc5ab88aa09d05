// tools.cpp -- TOOLS/*.f index algebra, descriptors, argument checks, PXERBLA; options & counters.
// Pure host integer code re-exported under the reference's Fortran symbols.
#include "common.h"

#include <cctype>
#include <map>
#include <mutex>

namespace slb {

// TOOLS/numroc.f
int numroc(int n, int nb, int iproc, int isrc, int nprocs)
{
    int mydist = (nprocs + iproc - isrc) % nprocs;
    int nblocks = n / nb;
    int r = (nblocks / nprocs) * nb;
    int extra = nblocks % nprocs;
    if (mydist < extra) r += nb;
    else if (mydist == extra) r += n % nb;
    return r;
}

// TOOLS/infog2l.f
void infog2l(int gr, int gc, const int *desc, int nprow, int npcol, int myrow, int mycol, int *lr, int *lc,
             int *rsrc, int *csrc)
{
    int mb = desc[MB_], nb = desc[NB_];
    int grc = gr - 1, gcc = gc - 1;
    int rblk = grc / mb, cblk = gcc / nb;
    *rsrc = (rblk + desc[RSRC_]) % nprow;
    *csrc = (cblk + desc[CSRC_]) % npcol;
    *lr = (rblk / nprow + 1) * mb + 1;
    *lc = (cblk / npcol + 1) * nb + 1;
    if ((myrow + nprow - desc[RSRC_]) % nprow >= rblk % nprow) {
        if (myrow == *rsrc) *lr += grc % mb;
        *lr -= mb;
    }
    if ((mycol + npcol - desc[CSRC_]) % npcol >= cblk % npcol) {
        if (mycol == *csrc) *lc += gcc % nb;
        *lc -= nb;
    }
}

// TOOLS/chk1mat.f:92-171.  INFO encoding: -(100*i+j) for entry j of array argument i, -i for scalar i.
void chk1mat(int ma, int mapos0, int na, int napos0, int ia, int ja, const int *desc, int descpos0, int *info)
{
    const int DM = 100, BIG = DM * DM;
    int inf = *info;
    if (inf >= 0) inf = BIG; else if (inf < -DM) inf = -inf; else inf = -inf * DM;
    const int mapos = mapos0 * DM, napos = napos0 * DM, iapos = (descpos0 - 2) * DM, japos = (descpos0 - 1) * DM,
              dpos = descpos0 * DM;
    int nprow = -1, npcol = -1, myrow = -1, mycol = -1;
    blacs_gridinfo_(&desc[CTXT_], &nprow, &npcol, &myrow, &mycol);
    auto lower = [&](int v) { if (v < inf) inf = v; };
    if (desc[DTYPE_] != 1) lower(dpos + DTYPE_ + 1);
    else if (ma < 0) lower(mapos);
    else if (na < 0) lower(napos);
    else if (ia < 1) lower(iapos);
    else if (ja < 1) lower(japos);
    else if (desc[MB_] < 1) lower(dpos + MB_ + 1);
    else if (desc[NB_] < 1) lower(dpos + NB_ + 1);
    else if (desc[RSRC_] < 0 || desc[RSRC_] >= nprow) lower(dpos + RSRC_ + 1);
    else if (desc[CSRC_] < 0 || desc[CSRC_] >= npcol) lower(dpos + CSRC_ + 1);
    else if (desc[LLD_] < 1) lower(dpos + LLD_ + 1);
    else if (desc[LLD_] < numroc(desc[M_], desc[MB_], myrow, desc[RSRC_], nprow)) {
        if (numroc(desc[N_], desc[NB_], mycol, desc[CSRC_], npcol) > 0) lower(dpos + LLD_ + 1);
    }
    if (ma == 0 || na == 0) {
        if (desc[M_] < 0) lower(dpos + M_ + 1);
        if (desc[N_] < 0) lower(dpos + N_ + 1);
    } else {
        if (desc[M_] < 1) lower(dpos + M_ + 1);
        else if (desc[N_] < 1) lower(dpos + N_ + 1);
        else {
            if (ia > desc[M_]) lower(iapos);
            else if (ja > desc[N_]) lower(japos);
            else {
                if (ia + ma - 1 > desc[M_]) lower(mapos);
                if (ja + na - 1 > desc[N_]) lower(napos);
            }
        }
    }
    if (inf == BIG) inf = 0; else if (inf % DM == 0) inf = -inf / DM; else inf = -inf;
    *info = inf;
}

// TOOLS/pchkxmat.f:404-490 (GLOBCHK): process (0,0)'s values are the reference; any process whose
// value k differs lowers INFO to that argument's position; then INFO = min over the whole grid.
static void globchk(int ictxt, int n, const int *vals, const int *pos, int *inf)
{
    Grid *g = grid_of(ictxt);
    if (!g || !g->in_grid()) return;
    int np = g->nprow * g->npcol;
    if (np > 1) {
        std::vector<int> all((size_t)np * n);
        grid_allgather(g, 'A', vals, all.data(), sizeof(int) * (size_t)n);
        // index 0 of scope 'A' is process (0,0)
        for (int k = 0; k < n; ++k)
            if (vals[k] != all[k] && pos[k] < *inf) *inf = pos[k];
        *inf = grid_imin(g, 'A', *inf);
    }
}

// ---- options & counters ----------------------------------------------------------------------
static std::mutex g_optmu;
static std::map<std::string, int64_t> &opts() { static std::map<std::string, int64_t> m; return m; }
static std::map<std::string, int64_t> &ctrs() { static std::map<std::string, int64_t> m; return m; }

int64_t opt(const char *key, int64_t dflt)
{
    std::lock_guard<std::mutex> lk(g_optmu);
    auto it = opts().find(key);
    if (it != opts().end()) return it->second;
    std::string env = "SLB200_"; for (const char *p = key; *p; ++p) env += (char)toupper(*p);
    const char *e = getenv(env.c_str());
    int64_t v = e && *e ? atoll(e) : dflt;
    opts()[key] = v;
    return v;
}
void counter_add(const char *key, int64_t v)
{
    std::lock_guard<std::mutex> lk(g_optmu);
    ctrs()[key] += v;
}

}  // namespace slb

using namespace slb;

extern "C" {

int numroc_(const int *n, const int *nb, const int *iproc, const int *isrcproc, const int *nprocs)
{ return numroc(*n, *nb, *iproc, *isrcproc, *nprocs); }
int indxg2p_(const int *ig, const int *nb, const int *iproc, const int *isrc, const int *nprocs)
{ (void)iproc; return indxg2p(*ig, *nb, *isrc, *nprocs); }
int indxg2l_(const int *ig, const int *nb, const int *iproc, const int *isrc, const int *nprocs)
{ (void)iproc; (void)isrc; return indxg2l(*ig, *nb, *nprocs); }
int indxl2g_(const int *il, const int *nb, const int *iproc, const int *isrc, const int *nprocs)
{ return indxl2g(*il, *nb, *iproc, *isrc, *nprocs); }
void infog2l_(const int *gr, const int *gc, const int *desc, const int *nprow, const int *npcol, const int *myrow,
              const int *mycol, int *lr, int *lc, int *rsrc, int *csrc)
{ infog2l(*gr, *gc, desc, *nprow, *npcol, *myrow, *mycol, lr, lc, rsrc, csrc); }
int iceil_(const int *a, const int *b) { return (*a + *b - 1) / *b; }   // TOOLS/iceil.f
int ilcm_(const int *m, const int *n)                                    // TOOLS/ilcm.f
{
    int ia = *m >= *n ? *m : *n, iq = *m >= *n ? *n : *m, ir;
    if (iq == 0) return 0;
    for (;;) { ir = ia % iq; if (ir == 0) break; ia = iq; iq = ir; }
    return (int)(((long long)*m * *n) / iq);
}

// PBLAS/SRC/PTZBLAS/pxerbla.f:53-58
void pxerbla_(const int *ictxt, const char *srname, const int *info)
{
    int nprow, npcol, myrow, mycol;
    blacs_gridinfo_(ictxt, &nprow, &npcol, &myrow, &mycol);
    char name[32]; int i = 0;
    // a Fortran CHARACTER has no terminator (its hidden length is not part of this C prototype): routine names are
    // [A-Z0-9_], at most 12 characters
    for (; i < 12 && (isalnum((unsigned char)srname[i]) || srname[i] == '_'); ++i) name[i] = srname[i];
    name[i] = 0;
    fprintf(stderr, "{%5d,%5d}:  On entry to %s parameter number%4d had an illegal value\n", myrow, mycol, name, *info);
    fflush(stderr);
}

// TOOLS/descinit.f:152-186
void descinit_(int *desc, const int *m, const int *n, const int *mb, const int *nb, const int *irsrc,
               const int *icsrc, const int *ictxt, const int *lld, int *info)
{
    int nprow, npcol, myrow, mycol;
    blacs_gridinfo_(ictxt, &nprow, &npcol, &myrow, &mycol);
    *info = 0;
    if (*m < 0) *info = -2;
    else if (*n < 0) *info = -3;
    else if (*mb < 1) *info = -4;
    else if (*nb < 1) *info = -5;
    else if (*irsrc < 0 || *irsrc >= nprow) *info = -6;
    else if (*icsrc < 0 || *icsrc >= npcol) *info = -7;
    else if (nprow == -1) *info = -8;
    else { int np = numroc(*m, *mb, myrow, *irsrc, nprow); if (*lld < (np > 1 ? np : 1)) *info = -9; }
    if (*info != 0) { int p = -*info; pxerbla_(ictxt, "DESCINIT", &p); }
    desc[DTYPE_] = 1;
    desc[M_] = *m > 0 ? *m : 0;
    desc[N_] = *n > 0 ? *n : 0;
    desc[MB_] = *mb > 1 ? *mb : 1;
    desc[NB_] = *nb > 1 ? *nb : 1;
    { int t = *irsrc < nprow - 1 ? *irsrc : nprow - 1; desc[RSRC_] = t > 0 ? t : 0; }
    { int t = *icsrc < npcol - 1 ? *icsrc : npcol - 1; desc[CSRC_] = t > 0 ? t : 0; }
    desc[CTXT_] = *ictxt;
    int np = nprow > 0 ? numroc(desc[M_], desc[MB_], myrow, desc[RSRC_], nprow) : 0;
    int t = np > 1 ? np : 1;
    desc[LLD_] = *lld > t ? *lld : t;
}

// TOOLS/descset.f
void descset_(int *desc, const int *m, const int *n, const int *mb, const int *nb, const int *irsrc,
              const int *icsrc, const int *ictxt, const int *lld)
{
    desc[DTYPE_] = 1; desc[CTXT_] = *ictxt; desc[M_] = *m; desc[N_] = *n; desc[MB_] = *mb; desc[NB_] = *nb;
    desc[RSRC_] = *irsrc; desc[CSRC_] = *icsrc; desc[LLD_] = *lld;
}

void chk1mat_(const int *ma, const int *mapos0, const int *na, const int *napos0, const int *ia, const int *ja,
              const int *desca, const int *descapos0, int *info)
{ chk1mat(*ma, *mapos0, *na, *napos0, *ia, *ja, desca, *descapos0, info); }

// TOOLS/pchkxmat.f:1-171
void pchk1mat_(const int *ma, const int *mapos0, const int *na, const int *napos0, const int *ia, const int *ja,
               const int *desca, const int *descapos0, const int *nextra, const int *ex, const int *expos, int *info)
{
    const int DM = 100, BIG = DM * DM;
    int inf = *info;
    if (inf >= 0) inf = BIG; else if (inf < -DM) inf = -inf; else inf = -inf * DM;
    int vals[64], pos[64], k = 0;
    auto put = [&](int v, int p) { vals[k] = v; pos[k] = p; ++k; };
    put(*ma, *mapos0 * DM); put(*na, *napos0 * DM); put(*ia, (*descapos0 - 2) * DM); put(*ja, (*descapos0 - 1) * DM);
    int dp = *descapos0 * DM;
    put(desca[DTYPE_], dp + 1); put(desca[M_], dp + 3); put(desca[N_], dp + 4); put(desca[MB_], dp + 5);
    put(desca[NB_], dp + 6); put(desca[RSRC_], dp + 7); put(desca[CSRC_], dp + 8);
    for (int i = 0; i < *nextra && k < 64; ++i) put(ex[i], expos[i]);
    globchk(desca[CTXT_], k, vals, pos, &inf);
    if (inf == BIG) inf = 0; else if (inf % DM == 0) inf = -inf / DM; else inf = -inf;
    *info = inf;
}

// TOOLS/pchkxmat.f:173-402
void pchk2mat_(const int *ma, const int *mapos0, const int *na, const int *napos0, const int *ia, const int *ja,
               const int *desca, const int *descapos0, const int *mb, const int *mbpos0, const int *nb,
               const int *nbpos0, const int *ib, const int *jb, const int *descb, const int *descbpos0,
               const int *nextra, const int *ex, const int *expos, int *info)
{
    const int DM = 100, BIG = DM * DM;
    int inf = *info;
    if (inf >= 0) inf = BIG; else if (inf < -DM) inf = -inf; else inf = -inf * DM;
    int vals[96], pos[96], k = 0;
    auto put = [&](int v, int p) { vals[k] = v; pos[k] = p; ++k; };
    put(*ma, *mapos0 * DM); put(*na, *napos0 * DM); put(*ia, (*descapos0 - 2) * DM); put(*ja, (*descapos0 - 1) * DM);
    int dp = *descapos0 * DM;
    put(desca[DTYPE_], dp + 1); put(desca[M_], dp + 3); put(desca[N_], dp + 4); put(desca[MB_], dp + 5);
    put(desca[NB_], dp + 6); put(desca[RSRC_], dp + 7); put(desca[CSRC_], dp + 8);
    put(*mb, *mbpos0 * DM); put(*nb, *nbpos0 * DM); put(*ib, (*descbpos0 - 2) * DM); put(*jb, (*descbpos0 - 1) * DM);
    dp = *descbpos0 * DM;
    put(descb[DTYPE_], dp + 1); put(descb[M_], dp + 3); put(descb[N_], dp + 4); put(descb[MB_], dp + 5);
    put(descb[NB_], dp + 6); put(descb[RSRC_], dp + 7); put(descb[CSRC_], dp + 8);
    for (int i = 0; i < *nextra && k < 96; ++i) put(ex[i], expos[i]);
    globchk(desca[CTXT_], k, vals, pos, &inf);
    if (inf == BIG) inf = 0; else if (inf % DM == 0) inf = -inf / DM; else inf = -inf;
    *info = inf;
}

// PBLAS/SRC/PTOOLS/PB_Ctop.c:76-141: process-global topology characters.  The NCCL data path has no
// use for them; they are stored and returned so PB_TOPGET/PB_TOPSET round-trip like the reference.
static char g_top[2][3] = { { ' ', ' ', ' ' }, { ' ', ' ', ' ' } };
static int top_op(const char *op) { return (op[0] == 'B' || op[0] == 'b') ? 0 : 1; }
static int top_scope(const char *sc) { char c = sc[0] & ~0x20; return c == 'R' ? 0 : (c == 'C' ? 1 : 2); }
void pb_topget_(const int *ictxt, const char *op, const char *scope, char *top)
{ (void)ictxt; *top = g_top[top_op(op)][top_scope(scope)]; }
void pb_topset_(const int *ictxt, const char *op, const char *scope, const char *top)
{ (void)ictxt; g_top[top_op(op)][top_scope(scope)] = *top; }

const char *slb200_version(void) { return "scalapack_b200 0.1 (sm_100a)"; }

}  // extern "C"

extern "C" void slb200_set_option(const char *key, int64_t value)
{
    std::lock_guard<std::mutex> lk(slb::g_optmu);
    slb::opts()[key] = value;
}
extern "C" int64_t slb200_get_counter(const char *key)
{
    std::lock_guard<std::mutex> lk(slb::g_optmu);
    auto it = slb::ctrs().find(key);
    return it == slb::ctrs().end() ? 0 : it->second;
}
extern "C" void slb200_reset_counters(void)
{
    std::lock_guard<std::mutex> lk(slb::g_optmu);
    slb::ctrs().clear();
}
