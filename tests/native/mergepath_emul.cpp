// mergepath_emul.cpp - TEST INFRASTRUCTURE: exposes the plan-time partition of spmm_mergepath.cu
// (crp-spmm_b200/csrc/cuda/mergepath_build.hpp) to tests/test_mergepath_structure.py.
#include <cstring>
#include "mergepath_build.hpp"

extern "C" int mergepath_partition(const int m, const int *rowptr, const int items, int *desc, const int desc_cap, int *long_row, int *long_sptr, int *counts)
{
    crp_mergepath_host h;
    crp_mergepath_partition(m, rowptr, items, &h);
    counts[0] = (int) (h.desc.size() / 4);  counts[1] = (int) h.long_row.size();  counts[2] = h.nseg;
    if ((int) h.desc.size() > desc_cap) return -1;
    memcpy(desc, h.desc.data(), sizeof(int) * h.desc.size());
    memcpy(long_row, h.long_row.data(), sizeof(int) * h.long_row.size());
    memcpy(long_sptr, h.long_sptr.data(), sizeof(int) * h.long_sptr.size());
    return 0;
}
