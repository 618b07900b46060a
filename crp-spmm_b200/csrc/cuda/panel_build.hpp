// panel_build.hpp - host-side construction of the "B row panel" form of a row-grouped matrix
// (pure C++, no CUDA: spmm_panel.cu uploads what this builds; tests/native/panel_emul.cpp walks
// the same structure on the CPU to check the builder without a GPU).
//
// Input: the row-group decomposition of spmm_rowgroup.cu - groups of R consecutive rows with one
// shared, strictly increasing column list, stored as R x 1 column blocks.
//
// A TILE is a set of K groups that are multiplied by ONE thread block (one consumer warp per group).
// Groups of an FEM / stencil matrix that are neighbours in the mesh use almost the same B rows, so the
// tile's B rows are brought into shared memory ONCE - the sorted union of the K column lists, the
// tile's "B row panel" - and every consumer reads the rows it needs from there.  Which groups share a
// tile is decided by crp_panel_cluster_tiles (greedy, by column overlap: for a mesh numbered line by
// line it finds compact 2-D / 3-D patches, whose panels are 20 - 50 % smaller than those of K
// consecutive groups) or, without clustering, K consecutive groups each.  The panel is cut into CHUNKS of at most CR rows / EMAX blocks; a chunk is
// what one pipeline stage of the kernel holds: CR row slices (bulk-copied, one cp.async.bulk per row)
// plus one contiguous "meta" record with everything the consumers need for it:
//
//   int32  nrows, flags (1: first chunk of its tile, 2: last chunk, 4: stop record)
//   int32  grow[K]        first C row of each consumer's group (-1: no such group in this tile)
//   int32  eoff[K + 1]    entry range of each consumer inside this chunk
//   (pad to 16 B)
//   uint32 slot[ne + 4]   low 16 bits: row of the chunk the entry multiplies; high 16 bits: mask of the
//                         group rows that really have the entry (all R bits set for exact groups)
//   (pad to 16 B)
//   T      val[ne + 4][R] the R values of the entry
//   (pad to 16 B)
// The four extra entries are zero (slot 0, mask 0, values 0): the kernel's software pipeline prefetches up to four
// entries past a consumer's range without a bounds check and never uses what it fetched there.
//
// Entries of a consumer appear in ascending column order across the chunks of a tile, i.e. a row's
// products are accumulated in the same order as the CSR row stores them (bit-identical to the row-split
// and row-group kernels when the column ids are the original ones).
#ifndef CRP_PANEL_BUILD_HPP
#define CRP_PANEL_BUILD_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

struct crp_rowgroup_host
{
    int R = 0;
    std::vector<int> g_row;             // ngroups: first row of each group
    std::vector<int> g_ptr;             // ngroups + 1: block range
    std::vector<int> b_col;             // nblk: (virtual) column of each block
    std::vector<unsigned short> b_mask; // nblk: rows of the group that have the block (empty: all of them)
    std::vector<double> b_val;          // nblk * R (absent rows: 0.0, never multiplied)
};

enum { CRP_PANEL_FIRST = 1, CRP_PANEL_LAST = 2, CRP_PANEL_STOP = 4 };

struct crp_panel_chunk { int uo0; int nrows; unsigned mo16; unsigned mlen16; };   // one 16-byte descriptor per chunk

struct crp_panel_host
{
    int R = 0, K = 0, CR = 0, EMAX = 0;
    int ntiles = 0;
    long long nentries = 0;
    std::vector<int> tile_chunk_ptr;            // ntiles + 1
    std::vector<crp_panel_chunk> chunks;        // nchunks + 1 (the last one is the stop record); mo16 / mlen16 are filled by crp_panel_fill_meta
    std::vector<int> ucol;                      // union column ids of all chunks
    // per chunk, structure only (independent of the value type): entries as (consumer-major) lists
    std::vector<int> ent_ptr;                   // nchunks * (K + 1) + 1 ... see crp_panel_build_structure
    std::vector<unsigned> ent_slot;             // slot | mask << 16
    std::vector<int> ent_blk;                   // index of the block in the row-group arrays
    std::vector<int> chunk_flags;               // nchunks
    std::vector<int> chunk_tile;                // nchunks
    std::vector<int> tile_groups;               // ntiles * K: the groups of each tile (-1: none), ascending
    int nchunks() const { return (int) chunk_flags.size(); }
    size_t hdr_bytes() const { return (size_t) ((3 + 2 * K + 3) / 4) * 16; }
    size_t meta_max(size_t elem) const
    {
        return hdr_bytes() + ((((size_t) EMAX + 4) * 4 + 15) & ~(size_t) 15) + ((((size_t) EMAX + 4) * (size_t) R * elem + 15) & ~(size_t) 15);
    }
};

// tiles, chunks, union columns and entry lists (no values)
// Greedy tile formation by column overlap.  Seeds are taken in group order; a tile grows by the unassigned group with the
// largest score = sum over the tile's members of the number of columns it shares with that member (ties: lowest group id),
// which prefers compact patches to chains; a tile that runs out of overlapping candidates is filled with the next
// unassigned groups in order.  Work: sum over tiles of K * |column list| * (groups per column) - linear in the number of
// blocks for bounded-degree meshes.  order: ntiles * K group ids, -1 = no group (last tile only).
static inline void crp_panel_cluster_tiles(const crp_rowgroup_host &rg, const int K, const int ncols, std::vector<int> *order)
{
    const int ng = (int) rg.g_row.size();
    order->clear();
    order->reserve(((size_t) ng + K - 1) / K * K);
    // inverted index: column -> groups that have a block in it
    std::vector<int> iptr((size_t) ncols + 1, 0);
    for (size_t b = 0; b < rg.b_col.size(); b++) iptr[(size_t) rg.b_col[b] + 1]++;
    for (int c = 0; c < ncols; c++) iptr[(size_t) c + 1] += iptr[c];
    std::vector<int> igrp(rg.b_col.size()), fill(iptr.begin(), iptr.end() - 1);
    for (int g = 0; g < ng; g++)
        for (int p = rg.g_ptr[g]; p < rg.g_ptr[g + 1]; p++) igrp[(size_t) fill[rg.b_col[p]]++] = g;
    std::vector<char> assigned((size_t) ng, 0);
    std::vector<int> score((size_t) ng, 0), touched;
    int next_free = 0;
    for (int seed = 0; seed < ng; seed++)
    {
        if (assigned[seed]) continue;
        touched.clear();
        int member = seed, size = 0;
        while (true)
        {
            assigned[member] = 1;
            order->push_back(member);
            if (++size == K) break;
            for (int p = rg.g_ptr[member]; p < rg.g_ptr[member + 1]; p++)
            {
                const int c = rg.b_col[p];
                for (int q = iptr[c]; q < iptr[(size_t) c + 1]; q++)
                {
                    const int g = igrp[q];
                    if (assigned[g]) continue;
                    if (score[g]++ == 0) touched.push_back(g);
                }
            }
            int best = -1, best_score = 0;
            for (size_t i = 0; i < touched.size(); i++)
            {
                const int g = touched[i];
                if (assigned[g]) continue;
                if (score[g] > best_score || (score[g] == best_score && g < best)) { best = g; best_score = score[g]; }
            }
            if (best < 0)
            {
                if (next_free <= seed) next_free = seed + 1;
                while (next_free < ng && assigned[next_free]) next_free++;
                if (next_free >= ng) break;
                best = next_free;
            }
            member = best;
        }
        for (size_t i = 0; i < touched.size(); i++) score[touched[i]] = 0;
        while (size++ < K) order->push_back(-1);
    }
}

static inline void crp_panel_build_structure(const crp_rowgroup_host &rg, const int K, const int CR, const int EMAX, crp_panel_host *ph,
                                             const std::vector<int> *order = NULL)
{
    ph->R = rg.R;  ph->K = K;  ph->CR = CR;  ph->EMAX = EMAX;
    const int ng = (int) rg.g_row.size();
    ph->ntiles = (ng + K - 1) / K;
    // members of each tile: K consecutive groups unless an order (crp_panel_cluster_tiles) is given; within a tile the groups
    // are kept in ascending order so that consumer w of a tile always has the w-th smallest first row
    ph->tile_groups.assign((size_t) ph->ntiles * K, -1);
    for (int t = 0; t < ph->ntiles; t++)
    {
        int *tg = ph->tile_groups.data() + (size_t) t * K;
        if (order != NULL) { for (int w = 0; w < K; w++) tg[w] = (*order)[(size_t) t * K + w]; }
        else { for (int w = 0; w < K; w++) tg[w] = (t * K + w < ng) ? t * K + w : -1; }
        std::sort(tg, tg + K, [](const int a, const int b) { return (a < 0) ? false : (b < 0 ? true : a < b); });
    }
    ph->tile_chunk_ptr.assign(1, 0);
    ph->chunks.clear();  ph->ucol.clear();  ph->ent_ptr.clear();  ph->ent_slot.clear();  ph->ent_blk.clear();
    ph->chunk_flags.clear();  ph->chunk_tile.clear();
    ph->nentries = 0;
    const unsigned full_mask = (1u << rg.R) - 1u;
    std::vector<int> uni, cur(K), cnt;
    std::vector<int> pend_(K);
    for (int t = 0; t < ph->ntiles; t++)
    {
        const int *tg = ph->tile_groups.data() + (size_t) t * K;
        // sorted union of the groups' column lists
        uni.clear();
        for (int w = 0; w < K; w++)
            if (tg[w] >= 0) uni.insert(uni.end(), rg.b_col.begin() + rg.g_ptr[tg[w]], rg.b_col.begin() + rg.g_ptr[tg[w] + 1]);
        std::sort(uni.begin(), uni.end());
        uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
        // how many groups use each union row (to respect EMAX when closing a chunk)
        cnt.assign(uni.size(), 0);
        for (int w = 0; w < K; w++)
        {
            if (tg[w] < 0) continue;
            size_t u = 0;
            for (int p = rg.g_ptr[tg[w]]; p < rg.g_ptr[tg[w] + 1]; p++)
            {
                while (uni[u] != rg.b_col[p]) u++;
                cnt[u]++;
            }
        }
        for (int w = 0; w < K; w++) { cur[w] = (tg[w] >= 0) ? rg.g_ptr[tg[w]] : 0; pend_[w] = (tg[w] >= 0) ? rg.g_ptr[tg[w] + 1] : 0; }
        size_t u0 = 0;
        const int first_chunk = ph->nchunks();
        do {
            // rows [u0, u1) of the union form the next chunk
            size_t u1 = u0;
            int ne = 0;
            while (u1 < uni.size() && (int) (u1 - u0) < CR && ne + cnt[u1] <= EMAX) { ne += cnt[u1]; u1++; }
            if (u1 == u0 && u0 < uni.size()) { ne = cnt[u1]; u1++; }       // a single row used by more than EMAX groups cannot happen (K <= EMAX), kept for safety
            crp_panel_chunk ck;
            ck.uo0 = (int) ph->ucol.size();  ck.nrows = (int) (u1 - u0);  ck.mo16 = 0;  ck.mlen16 = 0;
            ph->chunks.push_back(ck);
            ph->ucol.insert(ph->ucol.end(), uni.begin() + u0, uni.begin() + u1);
            const int hi = (u1 < uni.size()) ? uni[u1] : 0x7fffffff;       // columns below `hi` belong to this chunk
            for (int w = 0; w < K; w++)
            {
                ph->ent_ptr.push_back((int) ph->ent_slot.size());
                if (tg[w] < 0) continue;
                const int pend = pend_[w];
                size_t u = u0;
                while (cur[w] < pend && rg.b_col[cur[w]] < hi)
                {
                    while (uni[u] != rg.b_col[cur[w]]) u++;
                    const unsigned mask = rg.b_mask.empty() ? full_mask : (unsigned) rg.b_mask[cur[w]];
                    ph->ent_slot.push_back((unsigned) (u - u0) | (mask << 16));
                    ph->ent_blk.push_back(cur[w]);
                    cur[w]++;
                }
            }
            ph->chunk_flags.push_back(0);
            ph->chunk_tile.push_back(t);
            u0 = u1;
        } while (u0 < uni.size());
        ph->chunk_flags[first_chunk] |= CRP_PANEL_FIRST;
        ph->chunk_flags.back() |= CRP_PANEL_LAST;
        ph->tile_chunk_ptr.push_back(ph->nchunks());
    }
    ph->ent_ptr.push_back((int) ph->ent_slot.size());
    ph->nentries = (long long) ph->ent_slot.size();
    crp_panel_chunk stop;
    stop.uo0 = (int) ph->ucol.size();  stop.nrows = 0;  stop.mo16 = 0;  stop.mlen16 = 0;
    ph->chunks.push_back(stop);
}

// the meta records of all chunks (+ the stop record) for value type T; fills chunks[].mo16 / mlen16
template <typename T>
static inline void crp_panel_fill_meta(const crp_rowgroup_host &rg, crp_panel_host *ph, std::vector<unsigned char> *meta)
{
    const int K = ph->K, R = ph->R, nch = ph->nchunks();
    const size_t hdr = ph->hdr_bytes();
    size_t total = 0;
    std::vector<size_t> off((size_t) nch + 2);
    for (int c = 0; c < nch; c++)
    {
        const size_t ne = (size_t) (ph->ent_ptr[(size_t) (c + 1) * K] - ph->ent_ptr[(size_t) c * K]);
        off[c] = total;
        total += hdr + (((ne + 4) * 4 + 15) & ~(size_t) 15) + (((ne + 4) * (size_t) R * sizeof(T) + 15) & ~(size_t) 15);
    }
    off[nch] = total;
    total += hdr;               // stop record: header only
    off[nch + 1] = total;
    meta->assign(total, 0);
    for (int c = 0; c <= nch; c++)
    {
        unsigned char *rec = meta->data() + off[c];
        int *h = reinterpret_cast<int *>(rec);
        ph->chunks[c].mo16 = (unsigned) (off[c] / 16);
        ph->chunks[c].mlen16 = (unsigned) ((off[c + 1] - off[c]) / 16);
        if (c == nch) { h[0] = 0; h[1] = CRP_PANEL_STOP; for (int w = 0; w < K; w++) h[2 + w] = -1; continue; }
        const int t = ph->chunk_tile[c];
        const int ebase = ph->ent_ptr[(size_t) c * K];
        const size_t ne = (size_t) (ph->ent_ptr[(size_t) (c + 1) * K] - ebase);
        h[0] = ph->chunks[c].nrows;
        h[1] = ph->chunk_flags[c];
        for (int w = 0; w < K; w++)
        {
            const int g = ph->tile_groups[(size_t) t * K + w];
            h[2 + w] = (g >= 0) ? rg.g_row[g] : -1;
        }
        for (int w = 0; w <= K; w++) h[2 + K + w] = ph->ent_ptr[(size_t) c * K + w] - ebase;
        unsigned *slot = reinterpret_cast<unsigned *>(rec + hdr);
        T *val = reinterpret_cast<T *>(rec + hdr + (((ne + 4) * 4 + 15) & ~(size_t) 15));
        for (size_t e = 0; e < ne; e++)
        {
            slot[e] = ph->ent_slot[(size_t) ebase + e];
            const double *src = rg.b_val.data() + (size_t) ph->ent_blk[(size_t) ebase + e] * R;
            for (int r = 0; r < R; r++) val[e * R + r] = (T) src[r];
        }
    }
}

#endif
