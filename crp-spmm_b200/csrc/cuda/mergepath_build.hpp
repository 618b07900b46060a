// mergepath_build.hpp - plan-time partition for the nnz-balanced kernel (spmm_mergepath.cu); pure C++,
// shared with the CPU structure test (tests/native/mergepath_emul.cpp).
//
// The merge path of a CSR matrix interleaves "row end" items and "nonzero" items; cutting it every ITEMS items
// gives every worker the same amount of work whatever the row lengths are.  Here a cut is snapped back to the start
// of the row it falls into unless that row alone has ITEMS or more items, so that
//   - a chunk is either a run of WHOLE rows with at most ITEMS items (nonzeros + row ends), or
//   - a SEGMENT of at most ITEMS nonzeros of one long row (its partial result goes to scratch row `slot`,
//     a fix-up pass adds the segments of each long row in ascending order - deterministic, no atomics).
// Descriptor (4 ints per chunk): first row, number of rows (> 0) or -(slot + 1) for a segment, first nonzero, one past the last.
#ifndef CRP_MERGEPATH_BUILD_HPP
#define CRP_MERGEPATH_BUILD_HPP

#include <vector>

enum { CRP_MP_ITEMS = 256 };

struct crp_mergepath_host
{
    std::vector<int> desc;          // 4 per chunk
    std::vector<int> long_row;      // rows that were cut into segments
    std::vector<int> long_sptr;     // long rows + 1: their segment (= scratch row) ranges
    int nseg = 0;
};

static inline void crp_mergepath_partition(const int m, const int *rowptr, const int items, crp_mergepath_host *h)
{
    h->desc.clear();  h->long_row.clear();  h->long_sptr.assign(1, 0);  h->nseg = 0;
    int r0 = 0, used = 0;           // open chunk: rows [r0, r) holding `used` items
    auto flush = [&](const int r) {
        if (r > r0)
        {
            h->desc.push_back(r0);  h->desc.push_back(r - r0);  h->desc.push_back(rowptr[r0]);  h->desc.push_back(rowptr[r]);
        }
        r0 = r;  used = 0;
    };
    for (int r = 0; r < m; r++)
    {
        const int len = rowptr[r + 1] - rowptr[r];
        if (len + 1 > items)
        {
            flush(r);
            h->long_row.push_back(r);
            for (int p = rowptr[r]; p < rowptr[r + 1]; p += items)
            {
                const int pe = (p + items < rowptr[r + 1]) ? p + items : rowptr[r + 1];
                h->desc.push_back(r);  h->desc.push_back(-(h->nseg + 1));  h->desc.push_back(p);  h->desc.push_back(pe);
                h->nseg++;
            }
            h->long_sptr.push_back(h->nseg);
            r0 = r + 1;
            continue;
        }
        if (used + len + 1 > items) flush(r);
        used += len + 1;
    }
    flush(m);
}

#endif
