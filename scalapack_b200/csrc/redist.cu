// redist.cu -- SURVEY 8(f) row 2: PDGEMR2D / PZGEMR2D (REDIST/SRC/pdgemr.c:230-728, pgemraux.c): copy sub(A) of a matrix
// distributed on one process grid into sub(B) of a matrix with ANY other block-cyclic distribution (other block sizes, other
// source processes, another grid / context).  It is the converter that lets a caller with NB = 64 arrays run the LU at NB = 512.
//
// The reference walks the block intersections ("scanD0") and exchanges one message per process pair in a caterpillar schedule.
// Here: the elements process s owns under A's layout AND process d owns under B's layout are always a Cartesian product
// rows(s_row, d_row) x cols(s_col, d_col).  The two index lists per peer are built on the host (O(M + N) integers), one gather
// kernel per peer packs its block contiguously, ONE grouped ncclSend / ncclRecv exchange moves everything over NVLink, and one
// scatter kernel per peer puts the received block where B's layout wants it.  Both sides enumerate rows and columns in
// increasing global order, so no index travels with the data.
#include "common.h"
#include "dist.h"
#include "kernels.cuh"
#include "launch.h"
#include "ncclw.h"

#include <algorithm>

namespace slb {

namespace {

// buf[kr + kc * nr] = A[ridx[kr] + cidx[kc] * lda]  (PACK)  or the reverse (!PACK); indices are local to the window at A.
// TRANS (unpacking a transposed copy): the block arrives in the SOURCE's order, so its row kr is a column of the destination:
// A[ridx[kc] + cidx[kr] * lda] = buf[kr + kc * nr], ridx / cidx being the destination's row / column lists.
template <typename T, bool PACK, bool TRANS>
__global__ void __launch_bounds__(256)
block_move_kernel(int64_t nr, int64_t nc, const int *__restrict__ ridx, const int *__restrict__ cidx, T *__restrict__ A, int64_t lda,
                  T *__restrict__ buf)
{
    const int64_t total = nr * nc;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kr = e % nr, kc = e / nr;
        const int64_t a = TRANS ? (int64_t)ridx[kc] + (int64_t)cidx[kr] * lda : (int64_t)ridx[kr] + (int64_t)cidx[kc] * lda;
        if (PACK) buf[e] = A[a]; else A[a] = buf[e];
    }
}

// what every process of the global context tells the others about one side (pdgemr.c:322-352 does this with an IGAMN2D)
struct SideInfo { int in, P, Q, r, c, mb, nb, rsrc, csrc, i0, j0, gm, gn; };

struct Side {
    int P = 0, Q = 0, mb = 0, nb = 0, rsrc = 0, csrc = 0, i0 = 0, j0 = 0;
    std::vector<int> pos;                                       // pos[r * Q + c] = position of process (r, c) in the global context
    int owner_row(int i) const { return indxg2p(i0 + i + 1, mb, rsrc, P); }
    int owner_col(int j) const { return indxg2p(j0 + j + 1, nb, csrc, Q); }
    int lrow(int i) const { return indxg2l(i0 + i + 1, mb, P) - 1; }
    int lcol(int j) const { return indxg2l(j0 + j + 1, nb, Q) - 1; }
};

Side resolve(const char *which, const std::vector<SideInfo> &all, int off, int m, int n)
{
    Side s;
    const SideInfo *ref = nullptr;
    for (size_t p = 0; p < all.size() / 2; ++p) {
        const SideInfo &x = all[2 * p + off];
        if (!x.in) continue;
        if (!ref) ref = &x;
        else if (x.P != ref->P || x.Q != ref->Q || x.mb != ref->mb || x.nb != ref->nb || x.rsrc != ref->rsrc || x.csrc != ref->csrc ||
                 x.i0 != ref->i0 || x.j0 != ref->j0 || x.gm != ref->gm || x.gn != ref->gn)
            fatal("PDGEMR2D: the processes of the grid of %s disagree on its descriptor or offsets", which);
    }
    if (!ref) fatal("xxGEMR2D: something wrong in the parameters: no process of the global context is in the grid of %s", which);
    s.P = ref->P; s.Q = ref->Q; s.mb = ref->mb; s.nb = ref->nb; s.rsrc = ref->rsrc; s.csrc = ref->csrc; s.i0 = ref->i0; s.j0 = ref->j0;
    if (s.i0 < 0 || s.j0 < 0 || s.i0 + m > ref->gm || s.j0 + n > ref->gn || s.mb < 1 || s.nb < 1)
        fatal("PDGEMR2D: sub(%s) = (%d:%d, %d:%d) does not fit the %d x %d matrix", which, s.i0 + 1, s.i0 + m, s.j0 + 1, s.j0 + n, ref->gm, ref->gn);
    s.pos.assign((size_t)s.P * s.Q, -1);
    for (size_t p = 0; p < all.size() / 2; ++p) {
        const SideInfo &x = all[2 * p + off];
        if (x.in) s.pos[(size_t)x.r * s.Q + x.c] = (int)p;
    }
    for (int v : s.pos) if (v < 0) fatal("PDGEMR2D: the global context does not contain every process of the grid of %s", which);
    return s;
}

// lists[q] = the indices k in [0, len) that `mine` owns on side X (as a row index if xrows, else as a column index) and that
// process row / column q owns on side Y (as a row index if yrows, else as a column index).  ymap (optional): index k of side X
// corresponds to index ymap[k] of side Y (a row permutation between source and destination); the lists are then ordered by
// `order[k]` (the index on the SOURCE side), so that sender and receiver enumerate a block in the same order.
void split_by_peer(int len, int mine, bool xrows, bool yrows, const Side &X, const Side &Y, std::vector<std::vector<int>> &lists,
                   const int *ymap = nullptr, bool order_by_y = false)
{
    lists.assign((size_t)(yrows ? Y.P : Y.Q), std::vector<int>());
    for (int k = 0; k < len; ++k) {
        if ((xrows ? X.owner_row(k) : X.owner_col(k)) != mine) continue;
        const int ky = ymap ? ymap[k] : k;
        lists[(size_t)(yrows ? Y.owner_row(ky) : Y.owner_col(ky))].push_back(k);
    }
    if (ymap && order_by_y)
        for (auto &l : lists) std::sort(l.begin(), l.end(), [&](int a_, int b_) { return ymap[a_] < ymap[b_]; });
}

}  // namespace

// sub(B) <- sub(A) (tr = false; sub(A), sub(B) m x n) or sub(B) <- sub(A)^T (tr = true; sub(A) m x n, sub(B) n x m: PDTRAN's data
// movement, used by the PBLAS entry points of pblas.cu to bring op(A) into their working layout)
// rowmap (optional, tr = false only): row k of sub(B) <- row rowmap[k] of sub(A) (a row permutation on the way: PDLAPIV's movement)
template <typename T>
void gemr2d_core(int m, int n, const T *a, int ia, int ja, const int *desca, T *b, int ib, int jb, const int *descb, int gctxt, bool tr,
                 const int *rowmap)
{
    if (rowmap && tr) fatal("gemr2d_core: a row map cannot be combined with a transposition");
    std::vector<int> invmap;                                    // source row i goes to destination row invmap[i]
    if (rowmap) { invmap.resize((size_t)m); for (int k = 0; k < m; ++k) invmap[(size_t)rowmap[k]] = k; }
    if (m == 0 || n == 0) return;                               // pdgemr.c:303-304
    Grid *gg = grid_of(gctxt);
    if (!gg || !gg->in_grid()) return;                          // not a member of the global context: nothing to do, nothing to wait for
    const int np = gg->nprow * gg->npcol, me = gg->myrow * gg->npcol + gg->mycol;
    cudaStream_t s = rt().s_main;

    // ---- who is where: grid shapes, descriptors and coordinates of both sides, from every process of the global context ----
    SideInfo mine[2]; memset(mine, 0, sizeof(mine));
    const int *descs[2] = { desca, descb }; const int i0s[2] = { ia - 1, ib - 1 }, j0s[2] = { ja - 1, jb - 1 };
    Grid *gs[2] = { nullptr, nullptr };
    for (int k = 0; k < 2; ++k) {
        Grid *g = descs[k][CTXT_] >= 0 ? grid_of(descs[k][CTXT_]) : nullptr;
        if (!g || !g->in_grid()) continue;
        gs[k] = g;
        mine[k] = SideInfo{ 1, g->nprow, g->npcol, g->myrow, g->mycol, descs[k][MB_], descs[k][NB_], descs[k][RSRC_], descs[k][CSRC_], i0s[k], j0s[k],
                            descs[k][M_], descs[k][N_] };
    }
    std::vector<SideInfo> all((size_t)2 * np);
    if (np > 1) grid_allgather(gg, 'A', mine, all.data(), sizeof(mine)); else { all[0] = mine[0]; all[1] = mine[1]; }
    const Side SA = resolve("A", all, 0, m, n), SB = resolve("B", all, 1, tr ? n : m, tr ? m : n);

    // ---- what I send (as a process of A's grid) and what I receive (as a process of B's grid), peer by peer ----
    std::vector<size_t> scount((size_t)np, 0), sdispl((size_t)np, 0), rcount((size_t)np, 0), rdispl((size_t)np, 0);
    std::vector<std::vector<int>> srow, scol, rrow, rcol;
    // srow / rrow: lists over the ROW indices [0, m) of sub(A); scol / rcol: over its COLUMN indices [0, n).  Under tr a row
    // index of sub(A) is a column index of sub(B) and vice versa.
    // with a row map both sides order a block's rows by their SOURCE index: the sender's lists are over source rows anyway,
    // the receiver's lists are over destination rows k and are sorted by rowmap[k]
    if (gs[0]) { split_by_peer(m, gs[0]->myrow, true, !tr, SA, SB, srow, rowmap ? invmap.data() : nullptr, false); split_by_peer(n, gs[0]->mycol, false, tr, SA, SB, scol); }
    if (gs[1]) {
        split_by_peer(m, tr ? gs[1]->mycol : gs[1]->myrow, !tr, true, SB, SA, rrow, rowmap, true);
        split_by_peer(n, tr ? gs[1]->myrow : gs[1]->mycol, tr, false, SB, SA, rcol);
    }
    size_t stot = 0, rtot = 0;
    for (int p = 0; p < np; ++p) {                              // buffers are laid out in the order of the global context
        const SideInfo &pa = all[(size_t)2 * p], &pb = all[(size_t)2 * p + 1];
        if (gs[0] && pb.in) scount[(size_t)p] = srow[(size_t)(tr ? pb.c : pb.r)].size() * scol[(size_t)(tr ? pb.r : pb.c)].size() * sizeof(T);
        if (gs[1] && pa.in) rcount[(size_t)p] = rrow[(size_t)pa.r].size() * rcol[(size_t)pa.c].size() * sizeof(T);
        sdispl[(size_t)p] = stot; rdispl[(size_t)p] = rtot;
        stot += scount[(size_t)p]; rtot += rcount[(size_t)p];
    }

    // ---- index lists on the device: local positions inside the windows of sub(A) / sub(B) ----
    AnyWindow wa, wb; memset(&wa, 0, sizeof(wa)); memset(&wb, 0, sizeof(wb));
    if (gs[0]) wa = any_window(m, n, ia, ja, desca, SA.P, SA.Q, gs[0]->myrow, gs[0]->mycol);
    if (gs[1]) wb = any_window(tr ? n : m, tr ? m : n, ib, jb, descb, SB.P, SB.Q, gs[1]->myrow, gs[1]->mycol);
    std::vector<int> idx;                                       // [send rows by r1 | send cols by c1 | recv rows by r0 | recv cols by c0]
    std::vector<size_t> o_srow, o_scol, o_rrow, o_rcol;
    auto append = [&](const std::vector<std::vector<int>> &lists, std::vector<size_t> &offs, const Side &S, bool rows, int64_t loff) {
        for (const auto &l : lists) {
            offs.push_back(idx.size());
            for (int k : l) idx.push_back((rows ? S.lrow(k) : S.lcol(k)) - (int)loff);
        }
    };
    append(srow, o_srow, SA, true, wa.loff_r); append(scol, o_scol, SA, false, wa.loff_c);
    append(rrow, o_rrow, SB, !tr, tr ? wb.loff_c : wb.loff_r); append(rcol, o_rcol, SB, tr, tr ? wb.loff_r : wb.loff_c);
    int *idx_dev = (int *)workspace("rd_idx", (idx.size() + 1) * sizeof(int));
    if (!idx.empty()) SLB_CUDA(cudaMemcpyAsync(idx_dev, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    char *sbuf = (char *)workspace("rd_send", stot + 16), *rbuf = (char *)workspace("rd_recv", rtot + 16);

    // ---- pack, exchange, unpack ----
    StageMat<T> A("rd_stage_src", gs[0] ? a : nullptr, gs[0] ? desca[LLD_] : 1, wa.loff_r, wa.loff_c, gs[0] ? wa.mloc : 0, gs[0] ? wa.nloc : 0);
    StageMat<T> B("rd_stage_dst", gs[1] ? b : nullptr, gs[1] ? descb[LLD_] : 1, wb.loff_r, wb.loff_c, gs[1] ? wb.mloc : 0, gs[1] ? wb.nloc : 0, false);
    for (int p = 0; p < np && gs[0]; ++p) {
        if (!scount[(size_t)p]) continue;
        const SideInfo &pb = all[(size_t)2 * p + 1];
        const int qr = tr ? pb.c : pb.r, qc = tr ? pb.r : pb.c;
        const int64_t nr = (int64_t)srow[(size_t)qr].size(), nc = (int64_t)scol[(size_t)qc].size();
        const unsigned grid = grid1d(nr * nc);
        SLB_LAUNCH((block_move_kernel<T, true, false>), grid, 256, s, nr, nc, idx_dev + o_srow[(size_t)qr], idx_dev + o_scol[(size_t)qc], A.dev, A.ld,
                   reinterpret_cast<T *>(sbuf + sdispl[(size_t)p]));
    }
    if (np > 1 && !gg->nccl) gg->nccl = nccl_create(gg);
    nccl_alltoallv(np > 1 ? gg->nccl->all : nullptr, np, me, sbuf, scount.data(), sdispl.data(), rbuf, rcount.data(), rdispl.data(), s);
    for (int p = 0; p < np && gs[1]; ++p) {
        if (!rcount[(size_t)p]) continue;
        const SideInfo &pa = all[(size_t)2 * p];
        const int64_t nr = (int64_t)rrow[(size_t)pa.r].size(), nc = (int64_t)rcol[(size_t)pa.c].size();
        const unsigned grid = grid1d(nr * nc);
        // tr: the destination's ROW list is the one over sub(A)'s column indices (rcol), its COLUMN list the one over row indices (rrow)
        if (tr) SLB_LAUNCH((block_move_kernel<T, false, true>), grid, 256, s, nr, nc, idx_dev + o_rcol[(size_t)pa.c], idx_dev + o_rrow[(size_t)pa.r], B.dev, B.ld,
                           reinterpret_cast<T *>(rbuf + rdispl[(size_t)p]));
        else SLB_LAUNCH((block_move_kernel<T, false, false>), grid, 256, s, nr, nc, idx_dev + o_rrow[(size_t)pa.r], idx_dev + o_rcol[(size_t)pa.c], B.dev, B.ld,
                        reinterpret_cast<T *>(rbuf + rdispl[(size_t)p]));
    }
    SLB_CUDA(cudaStreamSynchronize(s));
    B.download();
}
template void gemr2d_core<double>(int, int, const double *, int, int, const int *, double *, int, int, const int *, int, bool, const int *);
template void gemr2d_core<zcomplex>(int, int, const zcomplex *, int, int, const int *, zcomplex *, int, int, const int *, int, bool, const int *);

// The public entry walks the columns in chunks so that the pack / receive buffers stay bounded (option redist_chunk_mb per process,
// default 1024): a 64 GiB local array is redistributed with 2 GiB of scratch, not 128.  Every process of the global context computes
// the same chunking from the global sizes.
template <typename T>
static void gemr2d_impl(int m, int n, const T *a, int ia, int ja, const int *desca, T *b, int ib, int jb, const int *descb, int gctxt)
{
    if (m <= 0 || n <= 0) return;
    Grid *gg = grid_of(gctxt);
    const int64_t np = gg && gg->in_grid() ? (int64_t)gg->nprow * gg->npcol : 1;
    const int64_t budget = opt("redist_chunk_mb", 1024) << 20;
    int64_t nc = budget * np / ((int64_t)m * (int64_t)sizeof(T));
    if (nc < 1) nc = 1;
    if (nc >= n) { gemr2d_core<T>(m, n, a, ia, ja, desca, b, ib, jb, descb, gctxt, false, nullptr); return; }
    for (int64_t j0 = 0; j0 < n; j0 += nc) {
        const int w = (int)(n - j0 < nc ? n - j0 : nc);
        gemr2d_core<T>(m, w, a, ia, ja + (int)j0, desca, b, ib, jb + (int)j0, descb, gctxt, false, nullptr);
    }
}

}  // namespace slb

using namespace slb;

extern "C" {

void pdgemr2d_(const int *m, const int *n, const double *a, const int *ia, const int *ja, const int *desca, double *b, const int *ib,
               const int *jb, const int *descb, const int *ictxt)
{ gemr2d_impl<double>(*m, *n, a, *ia, *ja, desca, b, *ib, *jb, descb, *ictxt); }
void pzgemr2d_(const int *m, const int *n, const slb200_z *a, const int *ia, const int *ja, const int *desca, slb200_z *b, const int *ib,
               const int *jb, const int *descb, const int *ictxt)
{ gemr2d_impl<zcomplex>(*m, *n, reinterpret_cast<const zcomplex *>(a), *ia, *ja, desca, reinterpret_cast<zcomplex *>(b), *ib, *jb, descb, *ictxt); }
void Cpdgemr2d(int m, int n, const double *a, int ia, int ja, const int *desca, double *b, int ib, int jb, const int *descb, int gcontext)
{ gemr2d_impl<double>(m, n, a, ia, ja, desca, b, ib, jb, descb, gcontext); }
void Cpzgemr2d(int m, int n, const slb200_z *a, int ia, int ja, const int *desca, slb200_z *b, int ib, int jb, const int *descb, int gcontext)
{ gemr2d_impl<zcomplex>(m, n, reinterpret_cast<const zcomplex *>(a), ia, ja, desca, reinterpret_cast<zcomplex *>(b), ib, jb, descb, gcontext); }

}  // extern "C"
