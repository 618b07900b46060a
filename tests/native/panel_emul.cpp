// panel_emul.cpp - TEST INFRASTRUCTURE, not part of the product: walks the structures that
// crp-spmm_b200/csrc/cuda/{rowgroup_build,panel_build}.hpp build for the sm_100a panel kernel on the
// CPU, decoding the very meta records the kernel's consumer warps decode, so that the builder and the
// record layout can be checked without a GPU (tests/test_panel_structure.py).  Built by the test with
//   g++ -O2 -ffp-contract=off -shared -fPIC -I crp-spmm_b200/csrc/cuda tests/native/panel_emul.cpp
// Products are accumulated with separate multiply and add, left to right, like the oracle's CSR loop
// (oracle/crp_oracle.c orc_csr_spmm), so results must be bit-identical to it.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rowgroup_build.hpp"

static int g_cluster = 0;
extern "C" void panel_emul_set_cluster(const int on) { g_cluster = on; }      // tiles by column overlap (crp_panel_cluster_tiles) or K consecutive groups

extern "C" int panel_emul_spmm(
    const int m, const int k, const int *rowptr, const int *colidx, const double *val, const int n,
    const double *B, double *C, const int K, const int CR, const int EMAX, const double min_fill, const int forced_R,
    long long *stats   /* [0] R, [1] ngroups, [2] nblk, [3] rest rows, [4] ntiles, [5] nchunks, [6] union rows, [7] relaxed blocks, [8] meta bytes, [9] max rows / chunk, [10] max entries / chunk */
)
{
    for (int i = 0; i < 11; i++) stats[i] = 0;
    crp_rg_rowinfo ri;
    crp_rg_scan_rows(m, rowptr, colidx, &ri);
    crp_rg_choice ch = crp_rg_choose(m, rowptr, colidx, ri, forced_R, min_fill, 0);
    std::vector<char> done((size_t) m, 0);
    stats[0] = ch.R;
    std::vector<int> rest;
    if (ch.R > 1)
    {
        crp_rowgroup_host rh;
        long long rest_nnz = 0;
        crp_rg_build(m, rowptr, colidx, val, ri, ch.R, ch.off, min_fill, &rh, &rest, &rest_nnz);
        crp_panel_host ph;
        std::vector<int> order;
        if (g_cluster) crp_panel_cluster_tiles(rh, K, k, &order);
        crp_panel_build_structure(rh, K, CR, EMAX, &ph, g_cluster ? &order : NULL);
        if (g_cluster)
        {
            // every group in exactly one tile
            std::vector<char> seen(rh.g_row.size(), 0);
            if (order.size() != (size_t) ph.ntiles * K) return -30;
            for (int g : order) { if (g < 0) continue; if (g >= (int) seen.size() || seen[g]) return -31; seen[g] = 1; }
            for (char c : seen) if (!c) return -32;
        }
        std::vector<unsigned char> meta;
        crp_panel_fill_meta<double>(rh, &ph, &meta);
        const int R = ch.R;
        stats[1] = (long long) rh.g_row.size();  stats[2] = (long long) rh.b_col.size();  stats[3] = (long long) rest.size();
        stats[4] = ph.ntiles;  stats[5] = ph.nchunks();  stats[6] = (long long) ph.ucol.size();  stats[8] = (long long) meta.size();
        for (size_t i = 0; i < rh.b_mask.size(); i++) if (rh.b_mask[i] != (1u << R) - 1u) stats[7]++;
        const size_t HDR = ph.hdr_bytes();
        if (meta.size() > ph.meta_max(8) * (size_t) (ph.nchunks() + 1)) return -10;
        std::vector<double> acc((size_t) K * R * n);
        std::vector<int> row0((size_t) K, -1);
        // the stream of one persistent block that owns every tile: chunk after chunk, then the stop record
        for (int c = 0; c <= ph.nchunks(); c++)
        {
            const crp_panel_chunk &ck = ph.chunks[c];
            if ((size_t) ck.mlen16 * 16 > ph.meta_max(8)) return -11;
            if (ck.nrows > CR) return -12;
            const unsigned char *rec = meta.data() + (size_t) ck.mo16 * 16;
            const int *hdr = (const int *) rec;
            if (hdr[1] & CRP_PANEL_STOP) { if (c != ph.nchunks()) return -13; break; }
            if (hdr[0] != ck.nrows) return -14;
            const int ne = hdr[2 + 2 * K];
            if (ne > EMAX && ck.nrows > 1) return -15;
            if (ck.nrows > stats[9]) stats[9] = ck.nrows;
            if (ne > stats[10]) stats[10] = ne;
            const unsigned *slots = (const unsigned *) (rec + HDR);
            const double *vals = (const double *) (rec + HDR + ((((size_t) ne + 4) * 4 + 15) & ~(size_t) 15));
            for (int e = ne; e < ne + 4; e++) { if (slots[e] != 0) return -22; for (int r = 0; r < R; r++) if (vals[(size_t) e * R + r] != 0.0) return -23; }
            const int *ucol = ph.ucol.data() + ck.uo0;
            for (int r = 1; r < ck.nrows; r++) if (ucol[r] <= ucol[r - 1]) return -16;     // panel rows strictly ascending
            for (int w = 0; w < K; w++)
            {
                if (hdr[1] & CRP_PANEL_FIRST)
                {
                    row0[w] = hdr[2 + w];
                    for (size_t i = 0; i < (size_t) R * n; i++) acc[(size_t) w * R * n + i] = 0.0;
                }
                for (int e = hdr[2 + K + w]; e < hdr[3 + K + w]; e++)
                {
                    const unsigned sm = slots[e];
                    const int slot = (int) (sm & 0xffffu);
                    if (slot >= ck.nrows) return -17;
                    const double *x = B + (size_t) ucol[slot] * n;
                    for (int r = 0; r < R; r++)
                    {
                        if (!((sm >> (16 + r)) & 1u)) { if (vals[(size_t) e * R + r] != 0.0) return -18; continue; }
                        double *a = acc.data() + ((size_t) w * R + r) * n;
                        const double v = vals[(size_t) e * R + r];
                        for (int j = 0; j < n; j++) a[j] = a[j] + v * x[j];
                    }
                }
                if ((hdr[1] & CRP_PANEL_LAST) && row0[w] >= 0)
                    for (int r = 0; r < R; r++)
                    {
                        if (done[(size_t) row0[w] + r]) return -19;
                        done[(size_t) row0[w] + r] = 1;
                        for (int j = 0; j < n; j++) C[((size_t) row0[w] + r) * n + j] = acc[((size_t) w * R + r) * n + j];
                    }
            }
        }
    } else {
        for (int i = 0; i < m; i++) rest.push_back(i);
        stats[3] = m;
    }
    for (int i : rest)
    {
        if (done[(size_t) i]) return -20;
        done[(size_t) i] = 1;
        for (int j = 0; j < n; j++) C[(size_t) i * n + j] = 0.0;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++)
            for (int j = 0; j < n; j++) C[(size_t) i * n + j] = C[(size_t) i * n + j] + val[p] * B[(size_t) colidx[p] * n + j];
    }
    for (int i = 0; i < m; i++) if (!done[(size_t) i]) return -21;
    return 0;
}
