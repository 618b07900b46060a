// rowgroup_build.hpp - plan-time row-group analysis of a CSR matrix (pure C++, no CUDA; shared by
// spmm_rowgroup.cu, spmm_panel.cu and the CPU structure tests in tests/native/).
//
// A group is R consecutive rows that are multiplied together: the union of their column lists is stored
// as R x 1 column blocks, one B-row load then feeds up to R FMAs per value.
//   exact group   : all R rows have the same, strictly increasing column list (FEM matrices with R unknowns
//                   per node, e.g. the 6 x 6 node blocks of pwtk) - every block is full;
//   relaxed group : the rows differ (boundary conditions, perturbed patterns); a block carries a mask of the
//                   rows that really have the column, absent rows hold 0.0 that is NEVER multiplied (so
//                   Inf / NaN in B behave as in the reference's CSR loop).  Accepted when at least
//                   `min_fill` of the R * |union| slots are real nonzeros.  Only the panel kernel runs these.
// Rows that end up in no group go to the row-split kernel through a row list.
#ifndef CRP_ROWGROUP_BUILD_HPP
#define CRP_ROWGROUP_BUILD_HPP

#include <algorithm>
#include <cstring>
#include <vector>

#include "panel_build.hpp"

static const int kCrpCandR[] = { 8, 6, 4, 3, 2 };

struct crp_rg_rowinfo
{
    std::vector<unsigned char> cont;    // row i has the same columns as row i - 1
    std::vector<unsigned char> incr;    // row i's columns are strictly increasing and the row is not empty
};

static inline void crp_rg_scan_rows(const int m, const int *rowptr, const int *colidx, crp_rg_rowinfo *ri)
{
    ri->cont.assign((size_t) m, 0);
    ri->incr.assign((size_t) m, 0);
    for (int i = 0; i < m; i++)
    {
        const int b = rowptr[i], len = rowptr[i + 1] - b;
        bool inc = len > 0;
        for (int p = b + 1; inc && p < b + len; p++) if (colidx[p] <= colidx[p - 1]) inc = false;
        ri->incr[(size_t) i] = inc;
        if (i > 0 && len > 0 && rowptr[i] - rowptr[i - 1] == len)
            ri->cont[(size_t) i] = (memcmp(colidx + b, colidx + rowptr[i - 1], sizeof(int) * (size_t) len) == 0);
    }
}

static inline bool crp_rg_exact(const crp_rg_rowinfo &ri, const int r0, const int R, const int m)
{
    if (r0 + R > m || !ri.incr[(size_t) r0]) return false;
    for (int r = r0 + 1; r < r0 + R; r++) if (!ri.cont[(size_t) r]) return false;
    return true;
}

// number of distinct columns of rows [r0, r0 + R) (all strictly increasing), -1 if some row is empty / unsorted
static inline int crp_rg_union_size(const crp_rg_rowinfo &ri, const int *rowptr, const int *colidx, const int r0, const int R, int *nnz_out)
{
    int pos[8], end[8], nnz = 0;
    for (int r = 0; r < R; r++)
    {
        if (!ri.incr[(size_t) (r0 + r)]) return -1;
        pos[r] = rowptr[r0 + r];  end[r] = rowptr[r0 + r + 1];
        nnz += end[r] - pos[r];
    }
    int nu = 0;
    for (;;)
    {
        int c = 0x7fffffff;
        for (int r = 0; r < R; r++) if (pos[r] < end[r] && colidx[pos[r]] < c) c = colidx[pos[r]];
        if (c == 0x7fffffff) break;
        for (int r = 0; r < R; r++) if (pos[r] < end[r] && colidx[pos[r]] == c) pos[r]++;
        nu++;
    }
    *nnz_out = nnz;
    return nu;
}

// is the group at r0 usable, and how many blocks does it make? (exact groups: the row length)
static inline bool crp_rg_group_ok(
    const crp_rg_rowinfo &ri, const int *rowptr, const int *colidx, const int r0, const int R, const int m, const double min_fill, int *nblk, bool *exact
)
{
    if (r0 + R > m) return false;
    if (crp_rg_exact(ri, r0, R, m)) { *nblk = rowptr[r0 + 1] - rowptr[r0]; *exact = true; return true; }
    if (min_fill >= 1.0) return false;
    int nnz = 0;
    const int nu = crp_rg_union_size(ri, rowptr, colidx, r0, R, &nnz);
    if (nu <= 0 || (double) nnz < min_fill * (double) R * (double) nu) return false;
    *nblk = nu;  *exact = false;
    return true;
}

struct crp_rg_choice { int R; int off; long long nblk; long long grouped_nnz; bool all_exact; };

// Cost unit: L1 / shared-memory wavefronts per 64 B of C row (4 per B-row load + 1 per row FMA'd); a group size must beat the
// row-split kernel (5 per nonzero) by 10 %.  Every alignment 0 .. R-1 of the first group is tried: a rank's first local row
// is generally not aligned with the matrix's node blocks.  `rows_limit` > 0 restricts the scan to the first rows (sampling).
static inline crp_rg_choice crp_rg_choose(
    const int m, const int *rowptr, const int *colidx, const crp_rg_rowinfo &ri, const int forced, const double min_fill, const int rows_limit
)
{
    const int mm = (rows_limit > 0 && rows_limit < m) ? rows_limit : m;
    const long long nnz_all = (long long) rowptr[mm] - rowptr[0];
    crp_rg_choice best = { 1, 0, 0, 0, true };
    double best_cost = (double) nnz_all * 5.0;
    for (int R : kCrpCandR)
    {
        if (forced > 1 && R != forced) continue;
        for (int off = 0; off < R && off < mm; off++)
        {
            long long nblk = 0, gnnz = 0;
            bool all_exact = true;
            for (int r0 = off; r0 + R <= mm; r0 += R)
            {
                int nb = 0;
                bool ex = true;
                if (!crp_rg_group_ok(ri, rowptr, colidx, r0, R, mm, min_fill, &nb, &ex)) continue;
                nblk += nb;
                gnnz += (long long) rowptr[r0 + R] - rowptr[r0];
                all_exact = all_exact && ex;
            }
            const double cost = (double) nblk * (4.0 + R) + (double) (nnz_all - gnnz) * 5.0;
            const bool take = (forced > 1) ? (nblk > 0 && (best.R == 1 || cost < best_cost)) : (cost < 0.9 * best_cost || (best.R == R && cost < best_cost));
            if (take) { best_cost = cost; best.R = R; best.off = off; best.nblk = nblk; best.grouped_nnz = gnnz; best.all_exact = all_exact; }
        }
    }
    return best;
}

// The decomposition for a given (R, off): group arrays + the rows left to the row-split kernel.
static inline void crp_rg_build(
    const int m, const int *rowptr, const int *colidx, const double *val, const crp_rg_rowinfo &ri, const int R, const int off,
    const double min_fill, crp_rowgroup_host *rh, std::vector<int> *rest_rows, long long *rest_nnz
)
{
    rh->R = R;
    rh->g_row.clear();  rh->g_ptr.assign(1, 0);  rh->b_col.clear();  rh->b_mask.clear();  rh->b_val.clear();
    rest_rows->clear();
    *rest_nnz = 0;
    bool any_relaxed = false;
    std::vector<unsigned short> masks;
    auto rest = [&](const int r) { rest_rows->push_back(r); *rest_nnz += rowptr[r + 1] - rowptr[r]; };
    for (int r = 0; r < off && r < m; r++) rest(r);
    for (int r0 = off; r0 < m; r0 += R)
    {
        const int r1 = std::min(m, r0 + R);
        int nb = 0;
        bool ex = true;
        if (!crp_rg_group_ok(ri, rowptr, colidx, r0, R, m, min_fill, &nb, &ex))
        {
            for (int r = r0; r < r1; r++) rest(r);
            continue;
        }
        rh->g_row.push_back(r0);
        if (ex)
        {
            const int len = rowptr[r0 + 1] - rowptr[r0];
            for (int j = 0; j < len; j++)
            {
                rh->b_col.push_back(colidx[rowptr[r0] + j]);
                masks.push_back((unsigned short) ((1u << R) - 1u));
                for (int r = 0; r < R; r++) rh->b_val.push_back(val[rowptr[r0 + r] + j]);
            }
        } else {
            any_relaxed = true;
            int pos[8], end[8];
            for (int r = 0; r < R; r++) { pos[r] = rowptr[r0 + r]; end[r] = rowptr[r0 + r + 1]; }
            for (;;)
            {
                int c = 0x7fffffff;
                for (int r = 0; r < R; r++) if (pos[r] < end[r] && colidx[pos[r]] < c) c = colidx[pos[r]];
                if (c == 0x7fffffff) break;
                unsigned mk = 0;
                rh->b_col.push_back(c);
                for (int r = 0; r < R; r++)
                {
                    if (pos[r] < end[r] && colidx[pos[r]] == c) { mk |= 1u << r; rh->b_val.push_back(val[pos[r]]); pos[r]++; }
                    else rh->b_val.push_back(0.0);
                }
                masks.push_back((unsigned short) mk);
            }
        }
        rh->g_ptr.push_back((int) rh->b_col.size());
    }
    if (any_relaxed) rh->b_mask.swap(masks);
}

#endif
