// Host test driver of the walk planner (extendablefem.jl_b200/csrc/walkplan.h): plain C ABI for ctypes.
#include "../extendablefem.jl_b200/csrc/walkplan.h"

extern "C" int tw_test_plan(int m, int L, const int *celloff, const int *kl, const int *orient, const int *pos /*[m][10]*/,
                            unsigned *words /*[m][12]*/, int *nload /* smem loads in first mode */, int *nstore)
{
    using namespace extfem;
    WalkTemplateIn T;
    T.m = m; T.L = L;
    for (int r = 0; r < m; ++r) {
        T.celloff.push_back(celloff[r]); T.kl.push_back(kl[r]); T.orient.push_back(orient[r]);
        std::array<int, 10> p;
        for (int t = 0; t < 10; ++t) p[t] = pos[r * 10 + t];
        T.pos.push_back(p);
    }
    std::vector<unsigned> w;
    if (!tw_plan_template(T, 1000, w)) return 0;
    int ld = 0, stc = 0;
    for (int r = 0; r < m; ++r) {
        const unsigned f = w[(size_t)r * TW_RW] & 0xffffu;
        for (int k = 0; k < TW_RW; ++k) words[r * TW_RW + k] = w[(size_t)r * TW_RW + k];
        if (f & TWF_EDGE) {
            ld += 3 * !!(f & TWF_LD0) + 3 * !!(f & TWF_LD1) + !!(f & TWF_LD2);
            stc += 3 + ((f & TWF_CO) ? 0 : 3) + 1;
        } else {
            ld += (f & TWF_WIN) ? 0 : 2 * !!(f & TWF_LD0) + 2 * !!(f & TWF_LD1);
            ld += 2 * !!(f & TWF_LD2) + !!(f & TWF_LDEI) + !!(f & TWF_LDEO) + !!(f & TWF_LDET);
            stc += 2 + 1 + ((f & TWF_CO) ? 0 : 1) + 1 + 2 * !!(f & TWF_FL1) + 2 * !!(f & TWF_FL2);
        }
    }
    *nload = ld; *nstore = stc;
    return 1;
}
