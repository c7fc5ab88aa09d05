/*
 * oracle.c -- CPU restatement of the reference ScaLAPACK dense-LU path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under scalapack_b200/ may include, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker
 * or as the timed CPU baseline.
 *
 * The reference itself (Fortran 77 + MPI) cannot be compiled in this image
 * (no gfortran, no MPI), so this file restates the algorithm the reference
 * executes, in the reference's own order, as a *serial* program acting on the
 * global matrix.  Data distribution is restated separately (orc_scatter /
 * orc_gather) so one process can emulate any P x Q block-cyclic layout.
 *
 * Parity pinning: the reference ships no stored numeric outputs for this
 * path.  The oracle is pinned by (1) the 6x6 PDGESV tutorial fixture
 * EXAMPLE/DSCAEXMAT.dat + DSCAEXRHS.dat (accept resid < 10,
 * EXAMPLE/pdscaex.f:181-192), (2) the LU.dat grid of cases with threshold 1.0
 * (TESTING/traditional/LU.dat:17), (3) an independent LAPACK dgetrf
 * (scipy) giving the same pivots, and -- since round 2 -- (4) the reference's
 * OWN Fortran, executed: tests/fortran77_mini.py interprets SRC/pdgetrf.f,
 * pdgetf2.f, pdlaswp.f, pdgetrs.f and the TOOLS index routines on a 1 x 1
 * grid (numpy for the PBLAS leaves); tests/test_reference_fortran.py holds
 * this file to their INFO / IPIV exactly and to their factors and solutions
 * to rounding (golden vectors in tests/golden/lu_reference.npz); the matrix
 * generator below is held BIT FOR BIT to TESTING/traditional/LIN/pdmatgen.f,
 * pzmatgen.f + pmatgeninc.f executed per process of P x Q grids
 * (tests/golden/matgen_reference.npz).  What stays
 * unpinned is the floating-point order inside the reference's external,
 * unversioned BLAS (CMakeLists.txt:160-190), which the reference itself does
 * not define.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { double re, im; } zdouble;

/* ------------------------------------------------------------------------- */
/* External BLAS (the reference binds these in PBLAS/SRC/PTOOLS/PB_Cdtypeset.c
 * :48-118); here: the OpenBLAS bundled with scipy, LP64, "scipy_" prefix.   */
/* ------------------------------------------------------------------------- */
typedef void (*dgemm_t)(const char*, const char*, const int*, const int*, const int*,
                        const double*, const double*, const int*, const double*, const int*,
                        const double*, double*, const int*);
typedef void (*dtrsm_t)(const char*, const char*, const char*, const char*, const int*, const int*,
                        const double*, const double*, const int*, double*, const int*);
typedef void (*dger_t)(const int*, const int*, const double*, const double*, const int*,
                       const double*, const int*, double*, const int*);
typedef int  (*idamax_t)(const int*, const double*, const int*);
typedef void (*dswap_t)(const int*, double*, const int*, double*, const int*);
typedef void (*dscal_t)(const int*, const double*, double*, const int*);
typedef void (*zscal_t)(const int*, const zdouble*, zdouble*, const int*);
typedef void (*setthr_t)(int);
typedef int  (*getthr_t)(void);

static struct {
    void *h;
    dgemm_t dgemm, zgemm;   /* zgemm has the same shape with zdouble pointers */
    dtrsm_t dtrsm, ztrsm;
    dger_t dger, zgeru;
    idamax_t idamax, izamax;
    dswap_t dswap, zswap;
    dscal_t dscal;
    zscal_t zscal;
    setthr_t setthr;
    getthr_t getthr;
} B;

int orc_init_blas(const char *path)
{
    if (B.h) return 0;
    B.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!B.h) return -1;
#define LD(field, name) *(void**)(&B.field) = dlsym(B.h, "scipy_" name)
    LD(dgemm, "dgemm_"); LD(zgemm, "zgemm_"); LD(dtrsm, "dtrsm_"); LD(ztrsm, "ztrsm_");
    LD(dger, "dger_"); LD(zgeru, "zgeru_"); LD(idamax, "idamax_"); LD(izamax, "izamax_");
    LD(dswap, "dswap_"); LD(zswap, "zswap_"); LD(dscal, "dscal_"); LD(zscal, "zscal_");
    LD(setthr, "openblas_set_num_threads"); LD(getthr, "openblas_get_num_threads");
#undef LD
    if (!B.dgemm || !B.dtrsm || !B.dger || !B.idamax || !B.dswap || !B.dscal) return -2;
    return 0;
}
void orc_set_threads(int n) { if (B.setthr) B.setthr(n); }
int  orc_get_threads(void) { return B.getthr ? B.getthr() : 1; }
int  orc_have_blas(void) { return B.h != NULL; }

double orc_wtime(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------------- */
/* TOOLS index algebra                                                       */
/* ------------------------------------------------------------------------- */
/* TOOLS/iceil.f */
int orc_iceil(int a, int b) { return (a + b - 1) / b; }

/* TOOLS/numroc.f:(whole file) */
int orc_numroc(int n, int nb, int iproc, int isrcproc, int nprocs)
{
    int mydist = (nprocs + iproc - isrcproc) % nprocs;
    int nblocks = n / nb;
    int r = (nblocks / nprocs) * nb;
    int extra = nblocks % nprocs;
    if (mydist < extra) r += nb;
    else if (mydist == extra) r += n % nb;
    return r;
}
/* TOOLS/indxg2p.f -- 1-based global index -> owning process coordinate */
int orc_indxg2p(int ig, int nb, int iproc, int isrc, int nprocs)
{ (void)iproc; return (isrc + (ig - 1) / nb) % nprocs; }
/* TOOLS/indxg2l.f -- 1-based global -> 1-based local */
int orc_indxg2l(int ig, int nb, int iproc, int isrc, int nprocs)
{ (void)iproc; (void)isrc; return nb * ((ig - 1) / (nb * nprocs)) + (ig - 1) % nb + 1; }
/* TOOLS/indxl2g.f -- 1-based local -> 1-based global */
int orc_indxl2g(int il, int nb, int iproc, int isrc, int nprocs)
{ return nprocs * nb * ((il - 1) / nb) + (il - 1) % nb + ((nprocs + iproc - isrc) % nprocs) * nb + 1; }

/* TOOLS/infog2l.f -- desc is the 9-int descriptor (DTYPE,CTXT,M,N,MB,NB,RSRC,CSRC,LLD) */
void orc_infog2l(int gr, int gc, const int *desc, int nprow, int npcol, int myrow, int mycol,
                 int *lr, int *lc, int *rsrc, int *csrc)
{
    int mb = desc[4], nb = desc[5], rs = desc[6], cs = desc[7];
    int grc = gr - 1, gcc = gc - 1;
    int rblk = grc / mb, cblk = gcc / nb;
    *rsrc = (rblk + rs) % nprow;
    *csrc = (cblk + cs) % npcol;
    *lr = (rblk / nprow + 1) * mb + 1;
    *lc = (cblk / npcol + 1) * nb + 1;
    if ((myrow + nprow - rs) % nprow >= rblk % nprow) {
        if (myrow == *rsrc) *lr += grc % mb;
        *lr -= mb;
    }
    if ((mycol + npcol - cs) % npcol >= cblk % npcol) {
        if (mycol == *csrc) *lc += gcc % nb;
        *lc -= nb;
    }
}

/* TOOLS/descinit.f:152-186 with the grid passed explicitly (no BLACS here).
 * Returns INFO; always fills desc with the clamped values like the reference. */
int orc_descinit(int *desc, int m, int n, int mb, int nb, int irsrc, int icsrc, int ictxt,
                 int lld, int nprow, int npcol, int myrow)
{
    int info = 0;
    if (m < 0) info = -2;
    else if (n < 0) info = -3;
    else if (mb < 1) info = -4;
    else if (nb < 1) info = -5;
    else if (irsrc < 0 || irsrc >= nprow) info = -6;
    else if (icsrc < 0 || icsrc >= npcol) info = -7;
    else if (nprow == -1) info = -8;
    else {
        int np = orc_numroc(m, mb, myrow, irsrc, nprow);
        if (lld < (np > 1 ? np : 1)) info = -9;
    }
    desc[0] = 1;
    desc[2] = m > 0 ? m : 0;
    desc[3] = n > 0 ? n : 0;
    desc[4] = mb > 1 ? mb : 1;
    desc[5] = nb > 1 ? nb : 1;
    { int t = irsrc < nprow - 1 ? irsrc : nprow - 1; desc[6] = t > 0 ? t : 0; }
    { int t = icsrc < npcol - 1 ? icsrc : npcol - 1; desc[7] = t > 0 ? t : 0; }
    desc[1] = ictxt;
    { int np = orc_numroc(desc[2], desc[4], myrow, desc[6], nprow);
      int t = np > 1 ? np : 1; desc[8] = lld > t ? lld : t; }
    return info;
}

/* TOOLS/chk1mat.f:92-171, grid passed explicitly.  info is in/out. */
void orc_chk1mat(int ma, int mapos0, int na, int napos0, int ia, int ja, const int *desc,
                 int descpos0, int nprow, int npcol, int myrow, int mycol, int *info)
{
    const int DM = 100, BIG = DM * DM;
    int inf = *info;
    if (inf >= 0) inf = BIG; else if (inf < -DM) inf = -inf; else inf = -inf * DM;
    int mapos = mapos0 * DM, napos = napos0 * DM, iapos = (descpos0 - 2) * DM,
        japos = (descpos0 - 1) * DM, dpos = descpos0 * DM;
#define MINI(x) do { if ((x) < inf) inf = (x); } while (0)
    if (desc[0] != 1) MINI(dpos + 1);
    else if (ma < 0) MINI(mapos);
    else if (na < 0) MINI(napos);
    else if (ia < 1) MINI(iapos);
    else if (ja < 1) MINI(japos);
    else if (desc[4] < 1) MINI(dpos + 5);
    else if (desc[5] < 1) MINI(dpos + 6);
    else if (desc[6] < 0 || desc[6] >= nprow) MINI(dpos + 7);
    else if (desc[7] < 0 || desc[7] >= npcol) MINI(dpos + 8);
    else if (desc[8] < 1) MINI(dpos + 9);
    else if (desc[8] < orc_numroc(desc[2], desc[4], myrow, desc[6], nprow)) {
        if (orc_numroc(desc[3], desc[5], mycol, desc[7], npcol) > 0) MINI(dpos + 9);
    }
    if (ma == 0 || na == 0) {
        if (desc[2] < 0) MINI(dpos + 3);
        if (desc[3] < 0) MINI(dpos + 4);
    } else {
        if (desc[2] < 1) MINI(dpos + 3);
        else if (desc[3] < 1) MINI(dpos + 4);
        else {
            if (ia > desc[2]) MINI(iapos);
            else if (ja > desc[3]) MINI(japos);
            else {
                if (ia + ma - 1 > desc[2]) MINI(mapos);
                if (ja + na - 1 > desc[3]) MINI(napos);
            }
        }
    }
#undef MINI
    if (inf == BIG) inf = 0; else if (inf % DM == 0) inf = -inf / DM; else inf = -inf;
    *info = inf;
}

/* ------------------------------------------------------------------------- */
/* Test-matrix generators                                                    */
/* ------------------------------------------------------------------------- */
/* The reference generator: X_{t+1} = (1103515245 X_t + 12345) mod 2^31
 * (TESTING/traditional/LIN/pdmatgen.f:117 MULT0/MULT1/IADD0, 16/15-bit limb
 * arithmetic pmatgeninc.f:5-78).  PDRAND returns X/2^31 then steps
 * (pmatgeninc.f:300-315).  */
#define LCG31_A 1103515245ULL
#define LCG31_C 12345ULL
#define LCG31_M 0x7fffffffULL
static inline uint64_t lcg31_step(uint64_t x) { return (LCG31_A * x + LCG31_C) & LCG31_M; }

/* a^k, c*(a^k-1)/(a-1) mod 2^31 by repeated squaring (what XJUMPM computes
 * by k-1 multiplications, pmatgeninc.f:80-130). */
static void lcg31_jump(uint64_t k, uint64_t *ak, uint64_t *ck)
{
    uint64_t a = LCG31_A, c = LCG31_C, ra = 1, rc = 0;
    while (k) {
        if (k & 1) { rc = (a * rc + c) & LCG31_M; ra = (a * ra) & LCG31_M; }
        c = ((a + 1) * c) & LCG31_M; a = (a * a) & LCG31_M;
        k >>= 1;
    }
    *ak = ra; *ck = rc;
}

/* Closed form of PDMATGEN 'N','N' (pdmatgen.f:448-510): global element (i,j),
 * 0-based, of an M x N matrix is 1 - 2*X_{1+i+j*M}/2^31, independent of grid
 * and block size.  Fills the full global matrix, column major. */
void orc_pdmatgen_global(int m, int n, int iseed, double *a, int64_t lda)
{
    uint64_t ak, ck, x1;
    lcg31_jump(1, &ak, &ck);
    x1 = (ak * (uint64_t)iseed + ck) & LCG31_M;          /* X_1 (JUMP1 = 1) */
    /* column j starts at X_{1+j*M}: jump by M per column (JUMP3 = M) */
    uint64_t am, cm; lcg31_jump((uint64_t)m, &am, &cm);
    uint64_t xc = x1;
    for (int j = 0; j < n; ++j) {
        uint64_t x = xc;
        for (int i = 0; i < m; ++i) {
            a[i + (int64_t)j * lda] = 1.0 - 2.0 * ((double)x / 2147483648.0);
            x = lcg31_step(x);
        }
        xc = (am * xc + cm) & LCG31_M;
    }
}

/* Structural restatement of PDMATGEN 'N','N' for one process of a P x Q grid
 * (pdmatgen.f:448-505): walks local blocks with the same seven jumps the
 * reference uses (JUMP1..JUMP7) instead of the closed form.  irsrc/icsrc are
 * IAROW/IACOL; generates the full local piece (IROFF=ICOFF=0). */
void orc_pdmatgen_local(int m, int n, int mb, int nb, double *a, int lda, int iarow, int iacol,
                        int iseed, int myrow, int mycol, int nprow, int npcol)
{
    int mp = orc_numroc(m, mb, myrow, iarow, nprow);
    int nq = orc_numroc(n, nb, mycol, iacol, npcol);
    int mrrow = (nprow + myrow - iarow) % nprow, mrcol = (npcol + mycol - iacol) % npcol;
    uint64_t a1, c1, a2, c2, a3, c3, a4, c4, a5, c5, t_a, t_c;
    lcg31_jump(1, &a1, &c1);
    lcg31_jump((uint64_t)nprow * mb, &a2, &c2);         /* JUMP2 = NPMB   */
    lcg31_jump((uint64_t)m, &a3, &c3);                  /* JUMP3 = M      */
    lcg31_jump((uint64_t)m * npcol * nb, &a4, &c4);     /* JUMP4 = NQNB columns */
    lcg31_jump((uint64_t)m * nb, &a5, &c5);             /* JUMP5 = NB columns   */
    uint64_t x = (a1 * (uint64_t)iseed + c1) & LCG31_M; /* IRAN1 after JUMP1    */
    lcg31_jump((uint64_t)m * nb * mrcol, &t_a, &t_c);   /* JUMP6 = MRCOL col blocks */
    x = (t_a * x + t_c) & LCG31_M;
    lcg31_jump((uint64_t)mb * mrrow, &t_a, &t_c);       /* JUMP7 = MB*MRROW rows    */
    x = (t_a * x + t_c) & LCG31_M;
    uint64_t ib1 = x, ib2 = x, ib3 = x;
    int jk = 0;
    int nend = orc_iceil(nq, nb), mend = orc_iceil(mp, mb);
    for (int ic = 0; ic < nend; ++ic) {
        for (int i = 0; i < nb; ++i) {
            if (jk >= nq) return;
            int ik = 0;
            for (int ir = 0; ir < mend; ++ir) {
                uint64_t r = ib1;
                for (int j = 0; j < mb; ++j) {
                    if (ik >= mp) break;
                    a[ik + (int64_t)jk * lda] = 1.0 - 2.0 * ((double)r / 2147483648.0);
                    r = lcg31_step(r);
                    ++ik;
                }
                if (ik >= mp) break;
                ib1 = (a2 * ib1 + c2) & LCG31_M;        /* JUMPIT(IA2,IC2,IB1) */
            }
            ++jk;
            ib2 = (a3 * ib2 + c3) & LCG31_M; ib1 = ib2; /* next column */
        }
        ib3 = (a4 * ib3 + c4) & LCG31_M; ib1 = ib2 = ib3; /* next local column block */
    }
    (void)a5; (void)c5;
}

/* Complex analogue (pzmatgen.f:475-511): all row jumps doubled,
 * a(i,j) = (1-2 X_{1+2(i+jM)}/2^31, 1-2 X_{2+2(i+jM)}/2^31). */
void orc_pzmatgen_global(int m, int n, int iseed, zdouble *a, int64_t lda)
{
    uint64_t ak, ck; lcg31_jump(1, &ak, &ck);
    uint64_t x1 = (ak * (uint64_t)iseed + ck) & LCG31_M;
    uint64_t am, cm; lcg31_jump(2ULL * (uint64_t)m, &am, &cm);
    uint64_t xc = x1;
    for (int j = 0; j < n; ++j) {
        uint64_t x = xc;
        for (int i = 0; i < m; ++i) {
            double re = 1.0 - 2.0 * ((double)x / 2147483648.0); x = lcg31_step(x);
            double im = 1.0 - 2.0 * ((double)x / 2147483648.0); x = lcg31_step(x);
            a[i + (int64_t)j * lda].re = re; a[i + (int64_t)j * lda].im = im;
        }
        xc = (am * xc + cm) & LCG31_M;
    }
}

/* 64-bit "HPL-style" generator for N beyond PDMATGEN's 2^31 period (NOT in the
 * reference): X_{t+1} = 6364136223846793005 X_t + 1 mod
 * 2^64, X_0 = seed; a(i,j) = (X_{1+i+j*M} >> 11) * 2^-53 - 0.5. */
#define LCG64_A 6364136223846793005ULL
#define LCG64_C 1ULL
static void lcg64_jump(uint64_t k, uint64_t *ak, uint64_t *ck)
{
    uint64_t a = LCG64_A, c = LCG64_C, ra = 1, rc = 0;
    while (k) {
        if (k & 1) { rc = a * rc + c; ra = a * ra; }
        c = (a + 1) * c; a = a * a;
        k >>= 1;
    }
    *ak = ra; *ck = rc;
}
static inline double lcg64_val(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5; }

/* rows [i0,i0+mr) x cols [j0,j0+nc) of the global M-row matrix into a(lda) */
void orc_matgen64_tile(int64_t m, uint64_t seed, int64_t i0, int64_t mr, int64_t j0, int64_t nc,
                       double *a, int64_t lda)
{
    for (int64_t j = 0; j < nc; ++j) {
        uint64_t ak, ck; lcg64_jump(1 + (uint64_t)i0 + (uint64_t)(j0 + j) * (uint64_t)m, &ak, &ck);
        uint64_t x = ak * seed + ck;
        for (int64_t i = 0; i < mr; ++i) { a[i + j * lda] = lcg64_val(x); x = LCG64_A * x + LCG64_C; }
    }
}
/* complex: element t = i + j*M uses X_{1+2t} (re) and X_{2+2t} (im) */
void orc_zmatgen64_tile(int64_t m, uint64_t seed, int64_t i0, int64_t mr, int64_t j0, int64_t nc,
                        zdouble *a, int64_t lda)
{
    for (int64_t j = 0; j < nc; ++j) {
        uint64_t ak, ck; lcg64_jump(1 + 2 * ((uint64_t)i0 + (uint64_t)(j0 + j) * (uint64_t)m), &ak, &ck);
        uint64_t x = ak * seed + ck;
        for (int64_t i = 0; i < mr; ++i) {
            a[i + j * lda].re = lcg64_val(x); x = LCG64_A * x + LCG64_C;
            a[i + j * lda].im = lcg64_val(x); x = LCG64_A * x + LCG64_C;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Block-cyclic scatter / gather (TOOLS/indxg2p.f, indxg2l.f applied per element
 * block); esz = element size in bytes (8 real, 16 complex).                  */
/* ------------------------------------------------------------------------- */
void orc_scatter(int m, int n, int mb, int nb, int rsrc, int csrc, int nprow, int npcol,
                 int myrow, int mycol, const void *ag, int64_t ldg, void *al, int64_t lld, int esz)
{
    const char *g = (const char*)ag; char *l = (char*)al;
    for (int j = 0; j < n; ++j) {
        if (orc_indxg2p(j + 1, nb, 0, csrc, npcol) != mycol) continue;
        int64_t jl = orc_indxg2l(j + 1, nb, 0, 0, npcol) - 1;
        for (int i0 = 0; i0 < m; i0 += mb) {
            if (orc_indxg2p(i0 + 1, mb, 0, rsrc, nprow) != myrow) continue;
            int ib = m - i0 < mb ? m - i0 : mb;
            int64_t il = orc_indxg2l(i0 + 1, mb, 0, 0, nprow) - 1;
            memcpy(l + (il + jl * lld) * esz, g + (i0 + (int64_t)j * ldg) * esz, (size_t)ib * esz);
        }
    }
}
void orc_gather(int m, int n, int mb, int nb, int rsrc, int csrc, int nprow, int npcol,
                int myrow, int mycol, void *ag, int64_t ldg, const void *al, int64_t lld, int esz)
{
    char *g = (char*)ag; const char *l = (const char*)al;
    for (int j = 0; j < n; ++j) {
        if (orc_indxg2p(j + 1, nb, 0, csrc, npcol) != mycol) continue;
        int64_t jl = orc_indxg2l(j + 1, nb, 0, 0, npcol) - 1;
        for (int i0 = 0; i0 < m; i0 += mb) {
            if (orc_indxg2p(i0 + 1, mb, 0, rsrc, nprow) != myrow) continue;
            int ib = m - i0 < mb ? m - i0 : mb;
            int64_t il = orc_indxg2l(i0 + 1, mb, 0, 0, nprow) - 1;
            memcpy(g + (i0 + (int64_t)j * ldg) * esz, l + (il + jl * lld) * esz, (size_t)ib * esz);
        }
    }
}

/* Local IPIV as the reference leaves it (SRC/pdgetrf.f:118-121): for process
 * row myrow, entry for each owned global row g < min(M,N) is the global pivot
 * ipiv_g[g]; entries for rows >= min(M,N) and the +MB spare are left at
 * `fill` (undefined in the reference: SURVEY.md section 8a note iv). */
void orc_ipiv_local(int m, int mn, int mb, int rsrc, int nprow, int myrow,
                    const int *ipiv_g, int *ipiv_l, int nloc, int fill)
{
    for (int i = 0; i < nloc; ++i) ipiv_l[i] = fill;
    for (int g = 0; g < mn && g < m; ++g) {
        if (orc_indxg2p(g + 1, mb, 0, rsrc, nprow) != myrow) continue;
        int il = orc_indxg2l(g + 1, mb, 0, 0, nprow) - 1;
        if (il < nloc) ipiv_l[il] = ipiv_g[g];
    }
}

/* ------------------------------------------------------------------------- */
/* Real LU: PDGETRF / PDGETF2 / PDLASWP / PDTRSM / PDGEMM restated serially   */
/* ------------------------------------------------------------------------- */
static inline int imin(int a, int b) { return a < b ? a : b; }

/* first index maximising |x| (BLAS idamax rule used by pdamax_.c:421-422;
 * across processes the combine keeps the earlier candidate on ties,
 * pdamax_.c:457, and candidates are ordered by process row = by global index
 * only within a block, so "first global index" is the serial equivalent on
 * tie-free input). */
static int idamax0(int n, const double *x)
{
    if (B.idamax) { int one = 1; return B.idamax(&n, x, &one) - 1; }
    int p = 0; double v = fabs(x[0]);
    for (int i = 1; i < n; ++i) if (fabs(x[i]) > v) { v = fabs(x[i]); p = i; }
    return p;
}

/* EXACT TIES.  With equal maxima the reference's answer depends on the process grid: PDAMAX combines the local candidates up a binary
 * tree towards process row 0 and the receiver keeps its own on a tie (strict <, pdamax_.c:436-458), so the candidate of the LOWEST
 * ABSOLUTE PROCESS ROW wins, and inside a process row the first local index (idamax).  On one process row that is the first global index
 * (the default here).  orc_set_tie_grid(nprow, rsrc) makes the serial restatement answer as an nprow-row grid whose process row rsrc owns
 * the first block row of sub(A) would: equilibrated matrices (PDGESVX, FACT = 'E') are full of entries that are exactly 1. */
static int g_tie_nprow = 1, g_tie_rsrc = 0;
void orc_set_tie_grid(int nprow, int rsrc) { g_tie_nprow = nprow > 1 ? nprow : 1; g_tie_rsrc = rsrc; }
static inline int tie_prow(int grow, int nb) { return (g_tie_rsrc + grow / nb) % g_tie_nprow; }

/* Unblocked panel: SRC/pdgetf2.f:207-237 on the m x jb panel at a (lda); row 0 of the panel is global row row0 of sub(A).
 * ipiv (0-based local to the panel) out; returns first zero-pivot column+1 or 0. */
static int getf2_ref(int m, int jb, double *a, int64_t lda_, int *ipiv, int row0, int nb)
{
    int info = 0, lda = (int)lda_, one = 1;
    int mn = imin(m, jb);
    for (int j = 0; j < mn; ++j) {
        int p = j + idamax0(m - j, a + j + (int64_t)j * lda);      /* PDAMAX  :212 */
        if (g_tie_nprow > 1) {
            const double v = fabs(a[p + (int64_t)j * lda]); int best = tie_prow(row0 + p, nb);
            for (int i = p + 1; i < m; ++i)
                if (fabs(a[i + (int64_t)j * lda]) == v && tie_prow(row0 + i, nb) < best) { best = tie_prow(row0 + i, nb); p = i; }
        }
        ipiv[j] = p;
        double gmax = a[p + (int64_t)j * lda];
        if (gmax != 0.0) {                                           /* :214 */
            if (p != j) {                                            /* PDSWAP over the jb panel columns :218 */
                if (B.dswap) B.dswap(&jb, a + j, &lda, a + p, &lda);
                else for (int c = 0; c < jb; ++c) { double t = a[j + (int64_t)c * lda]; a[j + (int64_t)c * lda] = a[p + (int64_t)c * lda]; a[p + (int64_t)c * lda] = t; }
            }
            if (j + 1 < m) {                                         /* PDSCAL by ONE/GMAX :224 */
                double r = 1.0 / gmax; int len = m - j - 1;
                if (B.dscal) B.dscal(&len, &r, a + j + 1 + (int64_t)j * lda, &one);
                else for (int i = j + 1; i < m; ++i) a[i + (int64_t)j * lda] *= r;
            }
        } else if (info == 0) info = j + 1;                          /* :226-227 */
        if (j + 1 < mn) {                                            /* PDGER :233 */
            int mm = m - j - 1, nn = jb - j - 1; double alpha = -1.0;
            if (B.dger) B.dger(&mm, &nn, &alpha, a + j + 1 + (int64_t)j * lda, &one,
                               a + j + (int64_t)(j + 1) * lda, &lda, a + j + 1 + (int64_t)(j + 1) * lda, &lda);
            else for (int c = j + 1; c < jb; ++c) { double u = a[j + (int64_t)c * lda];
                     for (int i = j + 1; i < m; ++i) a[i + (int64_t)c * lda] -= a[i + (int64_t)j * lda] * u; }
        }
    }
    return info;
}

/* SRC/pdlaswp.f:163-172 forward row interchanges k1..k2-1 (0-based, ipiv holds
 * 0-based global rows) over ncols columns starting at a. */
static void laswp_ref(int ncols, double *a, int64_t lda_, int k1, int k2, const int *ipiv)
{
    int lda = (int)lda_;
    if (ncols <= 0) return;
    for (int i = k1; i < k2; ++i) {
        int ip = ipiv[i];
        if (ip != i) {
            if (B.dswap && lda_ < 2147483647) B.dswap(&ncols, a + i, &lda, a + ip, &lda);
            else for (int c = 0; c < ncols; ++c) { double t = a[i + (int64_t)c * lda_]; a[i + (int64_t)c * lda_] = a[ip + (int64_t)c * lda_]; a[ip + (int64_t)c * lda_] = t; }
        }
    }
}

static void trsm_llnu_ref(int m, int n, const double *l, int64_t ldl, double *b, int64_t ldb)
{
    if (B.dtrsm) { double one = 1.0; int il = (int)ldl, ib = (int)ldb;
        B.dtrsm("L", "L", "N", "U", &m, &n, &one, l, &il, b, &ib); return; }
    for (int c = 0; c < n; ++c) for (int k = 0; k < m; ++k) { double x = b[k + c * ldb];
        for (int i = k + 1; i < m; ++i) b[i + c * ldb] -= l[i + k * ldl] * x; }
}
static void gemm_nn_minus_ref(int m, int n, int k, const double *a, int64_t lda, const double *b, int64_t ldb,
                              double *c, int64_t ldc)
{
    if (B.dgemm) { double mone = -1.0, one = 1.0; int ia = (int)lda, ib = (int)ldb, ic = (int)ldc;
        B.dgemm("N", "N", &m, &n, &k, &mone, a, &ia, b, &ib, &one, c, &ic); return; }
    for (int j = 0; j < n; ++j) for (int p = 0; p < k; ++p) { double u = b[p + j * ldb];
        for (int i = 0; i < m; ++i) c[i + j * ldc] -= a[i + p * lda] * u; }
}

/* SRC/pdgetrf.f:219-302 (IA=JA=1).  a: M x N global, column major, in place.
 * ipiv: min(M,N) entries, 1-based global row indices like the reference's
 * IPIV values (pdgetrf.f:118-121).  Returns INFO (0 or first zero pivot, 1-based).
 * t_panel/t_swap/t_trsm/t_gemm (nullable) accumulate phase wall times. */
int orc_dgetrf(int m, int n, double *a, int64_t lda, int nb, int *ipiv, double *phase_times)
{
    int mn = imin(m, n), info = 0;
    double tp = 0, ts = 0, tt = 0, tg = 0, t0;
    int *piv = (int*)malloc(sizeof(int) * (size_t)(mn > 0 ? mn : 1));
    for (int j0 = 0; j0 < mn; j0 += nb) {                              /* DO 10 :254 */
        int jb = imin(nb, mn - j0);
        t0 = orc_wtime();
        int iinfo = getf2_ref(m - j0, jb, a + j0 + (int64_t)j0 * lda, lda, piv + j0, j0, nb);   /* PDGETF2 :261 */
        for (int j = j0; j < j0 + imin(jb, m - j0); ++j) { piv[j] += j0; ipiv[j] = piv[j] + 1; }
        if (info == 0 && iinfo > 0) info = iinfo + j0;                 /* :263-264 */
        tp += orc_wtime() - t0; t0 = orc_wtime();
        laswp_ref(j0, a, lda, j0, j0 + jb, piv);                       /* PDLASWP left :268 */
        if (j0 + jb < n) {
            laswp_ref(n - j0 - jb, a + (int64_t)(j0 + jb) * lda, lda, j0, j0 + jb, piv);  /* right :275 */
            ts += orc_wtime() - t0; t0 = orc_wtime();
            trsm_llnu_ref(jb, n - j0 - jb, a + j0 + (int64_t)j0 * lda, lda,
                          a + j0 + (int64_t)(j0 + jb) * lda, lda);      /* PDTRSM :280 */
            tt += orc_wtime() - t0; t0 = orc_wtime();
            if (j0 + jb < m)
                gemm_nn_minus_ref(m - j0 - jb, n - j0 - jb, jb, a + j0 + jb + (int64_t)j0 * lda, lda,
                                  a + j0 + (int64_t)(j0 + jb) * lda, lda,
                                  a + j0 + jb + (int64_t)(j0 + jb) * lda, lda);           /* PDGEMM :288 */
            tg += orc_wtime() - t0;
        } else ts += orc_wtime() - t0;
    }
    free(piv);
    if (phase_times) { phase_times[0] = tp; phase_times[1] = ts; phase_times[2] = tt; phase_times[3] = tg; }
    return info;
}

/* Timed partial factorisation used as the bounded CPU-baseline sample: runs
 * only the first `nsteps` block steps of orc_dgetrf on an M x N matrix and
 * returns the flops those steps performed (panel + trsm + gemm) via *flops. */
int orc_dgetrf_steps(int m, int n, double *a, int64_t lda, int nb, int *ipiv, int nsteps, double *flops)
{
    int mn = imin(m, n), info = 0; double fl = 0;
    int *piv = (int*)malloc(sizeof(int) * (size_t)(mn > 0 ? mn : 1));
    int step = 0;
    for (int j0 = 0; j0 < mn && step < nsteps; j0 += nb, ++step) {
        int jb = imin(nb, mn - j0);
        int iinfo = getf2_ref(m - j0, jb, a + j0 + (int64_t)j0 * lda, lda, piv + j0, j0, nb);
        for (int j = j0; j < j0 + jb; ++j) { piv[j] += j0; ipiv[j] = piv[j] + 1; }
        if (info == 0 && iinfo > 0) info = iinfo + j0;
        double mm = m - j0, nn = n - j0 - jb;
        fl += mm * jb * jb - (double)jb * jb * jb / 3.0;   /* panel */
        laswp_ref(j0, a, lda, j0, j0 + jb, piv);
        if (j0 + jb < n) {
            laswp_ref(n - j0 - jb, a + (int64_t)(j0 + jb) * lda, lda, j0, j0 + jb, piv);
            trsm_llnu_ref(jb, n - j0 - jb, a + j0 + (int64_t)j0 * lda, lda, a + j0 + (int64_t)(j0 + jb) * lda, lda);
            fl += (double)jb * jb * nn;
            if (j0 + jb < m) {
                gemm_nn_minus_ref(m - j0 - jb, n - j0 - jb, jb, a + j0 + jb + (int64_t)j0 * lda, lda,
                                  a + j0 + (int64_t)(j0 + jb) * lda, lda, a + j0 + jb + (int64_t)(j0 + jb) * lda, lda);
                fl += 2.0 * (mm - jb) * nn * jb;
            }
        }
    }
    free(piv);
    if (flops) *flops = fl;
    return info;
}

/* SRC/pdgetrs.f:255-286.  a: factored N x N; ipiv 1-based global; b: N x nrhs in place.
 * trans: 'N' or 'T'/'C'. */
void orc_dgetrs(char trans, int n, int nrhs, const double *a, int64_t lda, const int *ipiv,
                double *b, int64_t ldb)
{
    int notran = (trans == 'N' || trans == 'n');
    if (n == 0 || nrhs == 0) return;
    if (notran) {
        for (int i = 0; i < n; ++i) {                                 /* PDLAPIV forward :255 */
            int p = ipiv[i] - 1;
            if (p != i) for (int c = 0; c < nrhs; ++c) { double t = b[i + c * ldb]; b[i + c * ldb] = b[p + c * ldb]; b[p + c * ldb] = t; }
        }
        for (int c = 0; c < nrhs; ++c) {                              /* PDTRSM L,L,N,U :260 */
            double *x = b + c * ldb;
            for (int k = 0; k < n; ++k) { double v = x[k]; if (v != 0.0) for (int i = k + 1; i < n; ++i) x[i] -= a[i + k * lda] * v; }
            for (int k = n - 1; k >= 0; --k) {                        /* PDTRSM L,U,N,N :265 */
                x[k] /= a[k + k * lda]; double v = x[k];
                for (int i = 0; i < k; ++i) x[i] -= a[i + k * lda] * v;
            }
        }
    } else {
        for (int c = 0; c < nrhs; ++c) {
            double *x = b + c * ldb;
            for (int k = 0; k < n; ++k) {                             /* U^T :273 */
                double s = x[k]; for (int i = 0; i < k; ++i) s -= a[i + k * lda] * x[i];
                x[k] = s / a[k + k * lda];
            }
            for (int k = n - 1; k >= 0; --k) {                        /* L^T unit :278 */
                double s = x[k]; for (int i = k + 1; i < n; ++i) s -= a[i + k * lda] * x[i];
                x[k] = s;
            }
        }
        for (int i = n - 1; i >= 0; --i) {                            /* PDLAPIV backward :283 */
            int p = ipiv[i] - 1;
            if (p != i) for (int c = 0; c < nrhs; ++c) { double t = b[i + c * ldb]; b[i + c * ldb] = b[p + c * ldb]; b[p + c * ldb] = t; }
        }
    }
}

/* infinity norm (PDLANGE 'I', SRC/pdlange.f) */
double orc_dlange_inf(int m, int n, const double *a, int64_t lda)
{
    double *s = (double*)calloc((size_t)(m > 0 ? m : 1), sizeof(double)), r = 0;
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) s[i] += fabs(a[i + j * lda]);
    for (int i = 0; i < m; ++i) if (s[i] > r || s[i] != s[i]) r = s[i];
    free(s); return r;
}

/* Factor residual: TESTING/traditional/LIN/pdgetrrv.f (rebuild P L U) then
 * pdlafchk.f:225-226: ||PLU - A||_inf / (max(M,N) eps ||A||_inf).
 * lu: factored M x N, a0: original.  eps = 2^-53 (PDLAMCH 'eps', SRC/pdlamch.f:68-78). */
double orc_fresid(int m, int n, const double *lu, int64_t ldlu, const int *ipiv, const double *a0, int64_t lda0)
{
    int mn = imin(m, n);
    double *r = (double*)calloc((size_t)m * n, sizeof(double));
    double *l = (double*)calloc((size_t)m * mn, sizeof(double));
    double *u = (double*)calloc((size_t)mn * n, sizeof(double));
    for (int j = 0; j < mn; ++j) { l[j + (int64_t)j * m] = 1.0; for (int i = j + 1; i < m; ++i) l[i + (int64_t)j * m] = lu[i + j * ldlu]; }
    for (int j = 0; j < n; ++j) for (int i = 0; i <= imin(j, mn - 1); ++i) u[i + (int64_t)j * mn] = lu[i + j * ldlu];
    if (B.dgemm) { double one = 1.0, zero = 0.0; B.dgemm("N", "N", &m, &n, &mn, &one, l, &m, u, &mn, &zero, r, &m); }
    else for (int j = 0; j < n; ++j) for (int p = 0; p < mn; ++p) { double v = u[p + (int64_t)j * mn]; for (int i = 0; i < m; ++i) r[i + (int64_t)j * m] += l[i + (int64_t)p * m] * v; }
    for (int i = mn - 1; i >= 0; --i) {                 /* undo interchanges: P applied backward */
        int p = ipiv[i] - 1;
        if (p != i) for (int c = 0; c < n; ++c) { double t = r[i + (int64_t)c * m]; r[i + (int64_t)c * m] = r[p + (int64_t)c * m]; r[p + (int64_t)c * m] = t; }
    }
    double anorm = orc_dlange_inf(m, n, a0, lda0);
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) r[i + (int64_t)j * m] -= a0[i + j * lda0];
    double rn = orc_dlange_inf(m, n, r, m);
    free(r); free(l); free(u);
    double eps = ldexp(1.0, -53);
    return rn / ((double)(m > n ? m : n) * eps * anorm);
}

/* Solve residual: TESTING/traditional/LIN/pdlaschk.f:187,296:
 * max_j ||b_j - A x_j||_inf / (||x_j||_inf ||A||_inf eps N). */
double orc_sresid(int n, int nrhs, const double *a0, int64_t lda0, const double *x, int64_t ldx,
                  const double *b0, int64_t ldb0)
{
    double anorm = orc_dlange_inf(n, n, a0, lda0), eps = ldexp(1.0, -53), res = 0;
    double *r = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int c = 0; c < nrhs; ++c) {
        for (int i = 0; i < n; ++i) r[i] = b0[i + c * ldb0];
        for (int j = 0; j < n; ++j) { double v = x[j + c * ldx]; for (int i = 0; i < n; ++i) r[i] -= a0[i + j * lda0] * v; }
        double rn = 0, xn = 0;
        for (int i = 0; i < n; ++i) { if (fabs(r[i]) > rn || r[i] != r[i]) rn = fabs(r[i]); if (fabs(x[i + c * ldx]) > xn) xn = fabs(x[i + c * ldx]); }
        double v = rn / (xn * anorm * eps * (double)n);
        if (v > res || v != v) res = v;
    }
    free(r); return res;
}

/* ------------------------------------------------------------------------- */
/* Complex LU: PZGETRF / PZGETF2 (SRC/pzgetrf.f, pzgetf2.f are textual type    */
/* swaps of the real files); pivot metric |Re|+|Im| (PBLAS/SRC/pzamax_.c:494). */
/* ------------------------------------------------------------------------- */
static inline double cabs1(zdouble z) { return fabs(z.re) + fabs(z.im); }
static inline zdouble zmul(zdouble a, zdouble b) { zdouble r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
/* complex reciprocal ONE/GMAX as the Fortran compiler evaluates (1,0)/z:
 * Smith's algorithm (what gfortran's __divdc3 does for finite inputs). */
static inline zdouble zrecip(zdouble z)
{
    zdouble r;
    if (fabs(z.re) >= fabs(z.im)) { double t = z.im / z.re, d = z.re + z.im * t; r.re = 1.0 / d; r.im = -t / d; }
    else { double t = z.re / z.im, d = z.re * t + z.im; r.re = t / d; r.im = -1.0 / d; }
    return r;
}

static int zgetf2_ref(int m, int jb, zdouble *a, int64_t lda, int *ipiv, int row0, int nb)
{
    int info = 0, mn = imin(m, jb);
    for (int j = 0; j < mn; ++j) {
        int p = j; double v = cabs1(a[j + j * lda]);
        for (int i = j + 1; i < m; ++i) { double w = cabs1(a[i + j * lda]); if (w > v) { v = w; p = i; } }
        if (g_tie_nprow > 1) {
            int best = tie_prow(row0 + p, nb);
            for (int i = p + 1; i < m; ++i)
                if (cabs1(a[i + j * lda]) == v && tie_prow(row0 + i, nb) < best) { best = tie_prow(row0 + i, nb); p = i; }
        }
        ipiv[j] = p;
        zdouble g = a[p + j * lda];
        if (g.re != 0.0 || g.im != 0.0) {
            if (p != j) for (int c = 0; c < jb; ++c) { zdouble t = a[j + c * lda]; a[j + c * lda] = a[p + c * lda]; a[p + c * lda] = t; }
            if (j + 1 < m) { zdouble r = zrecip(g); for (int i = j + 1; i < m; ++i) a[i + j * lda] = zmul(a[i + j * lda], r); }
        } else if (info == 0) info = j + 1;
        if (j + 1 < mn)
            for (int c = j + 1; c < jb; ++c) { zdouble u = a[j + c * lda];
                for (int i = j + 1; i < m; ++i) { zdouble t = zmul(a[i + j * lda], u); a[i + c * lda].re -= t.re; a[i + c * lda].im -= t.im; } }
    }
    return info;
}

int orc_zgetrf(int m, int n, zdouble *a, int64_t lda, int nb, int *ipiv)
{
    int mn = imin(m, n), info = 0;
    int *piv = (int*)malloc(sizeof(int) * (size_t)(mn > 0 ? mn : 1));
    for (int j0 = 0; j0 < mn; j0 += nb) {
        int jb = imin(nb, mn - j0);
        int iinfo = zgetf2_ref(m - j0, jb, a + j0 + j0 * lda, lda, piv + j0, j0, nb);
        for (int j = j0; j < j0 + jb; ++j) { piv[j] += j0; ipiv[j] = piv[j] + 1; }
        if (info == 0 && iinfo > 0) info = iinfo + j0;
        for (int i = j0; i < j0 + jb; ++i) { int ip = piv[i]; if (ip != i) {
            for (int c = 0; c < j0; ++c) { zdouble t = a[i + c * lda]; a[i + c * lda] = a[ip + c * lda]; a[ip + c * lda] = t; }
            for (int c = j0 + jb; c < n; ++c) { zdouble t = a[i + c * lda]; a[i + c * lda] = a[ip + c * lda]; a[ip + c * lda] = t; } } }
        if (j0 + jb < n) {
            int nn = n - j0 - jb;
            if (B.ztrsm) { zdouble one = { 1.0, 0.0 }; int il = (int)lda;
                B.ztrsm("L", "L", "N", "U", &jb, &nn, (double*)&one, (double*)(a + j0 + j0 * lda), &il, (double*)(a + j0 + (j0 + jb) * lda), &il);
            } else for (int c = j0 + jb; c < n; ++c) for (int k = j0; k < j0 + jb; ++k) { zdouble x = a[k + c * lda];
                    for (int i = k + 1; i < j0 + jb; ++i) { zdouble t = zmul(a[i + k * lda], x); a[i + c * lda].re -= t.re; a[i + c * lda].im -= t.im; } }
            if (j0 + jb < m) {
                int mm = m - j0 - jb;
                if (B.zgemm) { zdouble mone = { -1.0, 0.0 }, one = { 1.0, 0.0 }; int il = (int)lda;
                    B.zgemm("N", "N", &mm, &nn, &jb, (double*)&mone, (double*)(a + j0 + jb + j0 * lda), &il,
                            (double*)(a + j0 + (j0 + jb) * lda), &il, (double*)&one, (double*)(a + j0 + jb + (j0 + jb) * lda), &il);
                } else for (int c = j0 + jb; c < n; ++c) for (int k = j0; k < j0 + jb; ++k) { zdouble u = a[k + c * lda];
                        for (int i = j0 + jb; i < m; ++i) { zdouble t = zmul(a[i + k * lda], u); a[i + c * lda].re -= t.re; a[i + c * lda].im -= t.im; } }
            }
        }
    }
    free(piv);
    return info;
}

static inline zdouble zdiv(zdouble a, zdouble b) { return zmul(a, zrecip(b)); }

void orc_zgetrs(char trans, int n, int nrhs, const zdouble *a, int64_t lda, const int *ipiv, zdouble *b, int64_t ldb)
{
    if (n == 0 || nrhs == 0) return;
    if (!(trans == 'N' || trans == 'n')) {
        /* SRC/pzgetrs.f:268-284: op(U)^-1 (forward), op(L)^-1 unit (backward), PZLAPIV backward; op = ^T or ^H ('C') */
        const int cj = (trans == 'C' || trans == 'c');
        for (int c = 0; c < nrhs; ++c) {
            zdouble *x = b + c * ldb;
            for (int k = 0; k < n; ++k) {
                zdouble s = x[k];
                for (int i = 0; i < k; ++i) { zdouble e = a[i + k * lda]; if (cj) e.im = -e.im; zdouble t = zmul(e, x[i]); s.re -= t.re; s.im -= t.im; }
                zdouble d = a[k + k * lda]; if (cj) d.im = -d.im;
                x[k] = zdiv(s, d);
            }
            for (int k = n - 1; k >= 0; --k) {
                zdouble s = x[k];
                for (int i = k + 1; i < n; ++i) { zdouble e = a[i + k * lda]; if (cj) e.im = -e.im; zdouble t = zmul(e, x[i]); s.re -= t.re; s.im -= t.im; }
                x[k] = s;
            }
        }
        for (int i = n - 1; i >= 0; --i) { int p = ipiv[i] - 1;
            if (p != i) for (int c = 0; c < nrhs; ++c) { zdouble t = b[i + c * ldb]; b[i + c * ldb] = b[p + c * ldb]; b[p + c * ldb] = t; } }
        return;
    }
    for (int i = 0; i < n; ++i) { int p = ipiv[i] - 1;
        if (p != i) for (int c = 0; c < nrhs; ++c) { zdouble t = b[i + c * ldb]; b[i + c * ldb] = b[p + c * ldb]; b[p + c * ldb] = t; } }
    for (int c = 0; c < nrhs; ++c) {
        zdouble *x = b + c * ldb;
        for (int k = 0; k < n; ++k) { zdouble v = x[k]; for (int i = k + 1; i < n; ++i) { zdouble t = zmul(a[i + k * lda], v); x[i].re -= t.re; x[i].im -= t.im; } }
        for (int k = n - 1; k >= 0; --k) { x[k] = zdiv(x[k], a[k + k * lda]); zdouble v = x[k];
            for (int i = 0; i < k; ++i) { zdouble t = zmul(a[i + k * lda], v); x[i].re -= t.re; x[i].im -= t.im; } }
    }
}

/* PZLANGE 'I' uses the true modulus */
static double zlange_inf(int m, int n, const zdouble *a, int64_t lda)
{
    double *s = (double*)calloc((size_t)(m > 0 ? m : 1), sizeof(double)), r = 0;
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) s[i] += hypot(a[i + j * lda].re, a[i + j * lda].im);
    for (int i = 0; i < m; ++i) if (s[i] > r || s[i] != s[i]) r = s[i];
    free(s); return r;
}

double orc_zfresid(int m, int n, const zdouble *lu, int64_t ldlu, const int *ipiv, const zdouble *a0, int64_t lda0)
{
    int mn = imin(m, n);
    zdouble *r = (zdouble*)calloc((size_t)m * n, sizeof(zdouble));
    for (int j = 0; j < n; ++j)
        for (int p = 0; p <= imin(j, mn - 1); ++p) {
            zdouble u = lu[p + j * ldlu];
            r[p + (int64_t)j * m].re += u.re; r[p + (int64_t)j * m].im += u.im;   /* unit diagonal of L */
            for (int i = p + 1; i < m; ++i) { zdouble t = zmul(lu[i + p * ldlu], u); r[i + (int64_t)j * m].re += t.re; r[i + (int64_t)j * m].im += t.im; }
        }
    for (int i = mn - 1; i >= 0; --i) { int p = ipiv[i] - 1;
        if (p != i) for (int c = 0; c < n; ++c) { zdouble t = r[i + (int64_t)c * m]; r[i + (int64_t)c * m] = r[p + (int64_t)c * m]; r[p + (int64_t)c * m] = t; } }
    double anorm = zlange_inf(m, n, a0, lda0);
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) { r[i + (int64_t)j * m].re -= a0[i + j * lda0].re; r[i + (int64_t)j * m].im -= a0[i + j * lda0].im; }
    double rn = zlange_inf(m, n, r, m);
    free(r);
    return rn / ((double)(m > n ? m : n) * ldexp(1.0, -53) * anorm);
}

double orc_zsresid(int n, int nrhs, const zdouble *a0, int64_t lda0, const zdouble *x, int64_t ldx, const zdouble *b0, int64_t ldb0)
{
    double anorm = zlange_inf(n, n, a0, lda0), eps = ldexp(1.0, -53), res = 0;
    zdouble *r = (zdouble*)malloc(sizeof(zdouble) * (size_t)(n > 0 ? n : 1));
    for (int c = 0; c < nrhs; ++c) {
        for (int i = 0; i < n; ++i) r[i] = b0[i + c * ldb0];
        for (int j = 0; j < n; ++j) { zdouble v = x[j + c * ldx]; for (int i = 0; i < n; ++i) { zdouble t = zmul(a0[i + j * lda0], v); r[i].re -= t.re; r[i].im -= t.im; } }
        double rn = 0, xn = 0;
        for (int i = 0; i < n; ++i) { double w = hypot(r[i].re, r[i].im); if (w > rn || w != w) rn = w; w = hypot(x[i + c * ldx].re, x[i + c * ldx].im); if (w > xn) xn = w; }
        double v = rn / (xn * anorm * eps * (double)n);
        if (v > res || v != v) res = v;
    }
    free(r); return res;
}
