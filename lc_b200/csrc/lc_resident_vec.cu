// lc_b200 — vectorised instantiations of the resident kernel (lc_resident_kernel.cuh, lc_vec.cuh): planar 16-byte aligned
// fp32 slabs, four consecutive points per thread.  Host-side eligibility and dispatch: lc_resident.cu.
#include "lc_resident_kernel.cuh"

namespace lc {

int launch_res_vec(const lc_args& a, int mode, int nt, bool tm, cudaStream_t st, const ResLaunch& r) {
    return launch_res_any<true, 1>(a, mode, nt, tm, st, r);
}

}  // namespace lc
