// Mutating operations and tree export (Add_Points, Delete_Points, Delete_Point_Boxes, Update/refit,
// Criterion_Check + Rebuild, flatten). See the per-function comments for the reference lines replaced.
#include <algorithm>
#include <vector>

#include "ikd_host.h"

namespace ikd {

// Pre-order structure dump for parity tests (columns as oracle/ref_harness.cpp ref_dump_tree).
int dump_tree_impl(ikd_tree* t, float* out, int64_t cap, int64_t* out_n) {
    IKD_TRY(sync_header(t));
    *out_n = 0;
    if (!t->hdr.root_exists) return IKD_OK;
    size_t used = t->hdr.pool_top;
    std::vector<SearchRec> sr(used);
    std::vector<UpdateRec> ur(used);
    IKD_CUDA(cudaMemcpy(sr.data(), t->srec, used * sizeof(SearchRec), cudaMemcpyDeviceToHost));
    IKD_CUDA(cudaMemcpy(ur.data(), t->urec, used * sizeof(UpdateRec), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> stack;
    stack.push_back(ROOT_SLOT);
    int64_t k = 0;
    while (!stack.empty()) {
        uint32_t s = stack.back();
        stack.pop_back();
        const SearchRec& a = sr[s];
        const UpdateRec& u = ur[s];
        uint32_t cp = meta_cp(a.meta);
        bool hl = cp && (ur[2 * cp].flags & F_EXISTS), hr = cp && (ur[2 * cp + 1].flags & F_EXISTS);
        if (k < cap) {
            float* o = out + 16 * k;
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = (float)meta_axis(a.meta);
            o[4] = (float)u.size; o[5] = (float)u.invalid;
            o[6] = (float)(((u.flags & F_PDEL) ? 1 : 0) | ((u.flags & F_TDEL) ? 2 : 0) | ((u.flags & F_PDS) ? 4 : 0) |
                           ((u.flags & F_TDS) ? 8 : 0));
            o[7] = u.bmin[0]; o[8] = u.bmax[0]; o[9] = u.bmin[1]; o[10] = u.bmax[1]; o[11] = u.bmin[2]; o[12] = u.bmax[2];
            o[13] = hl ? 1.f : 0.f; o[14] = hr ? 1.f : 0.f; o[15] = (float)u.down_del;
        }
        k++;
        if (hr) stack.push_back(2 * cp + 1);
        if (hl) stack.push_back(2 * cp);
    }
    *out_n = k;
    return IKD_OK;
}

}  // namespace ikd
