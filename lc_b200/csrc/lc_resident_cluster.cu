// lc_b200 — cluster-split instantiations of the vectorised resident kernel (lc_resident_kernel.cuh, CL = 2): one pose on a
// thread-block cluster of two CTAs, each holding half of the points, reductions completed through distributed shared memory.
// lc_resident.cu uses them for the poses of the last, mostly empty wave of a launch and for batches smaller than half a wave.
#include "lc_resident_kernel.cuh"

namespace lc {

int launch_res_cluster2(const lc_args& a, int mode, int nt, cudaStream_t st, const ResLaunch& r) {
    return launch_res_any<true, 2>(a, mode, nt, false, st, r);
}

}  // namespace lc
