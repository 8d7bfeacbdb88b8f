"""Drop-in for the reference's compiled cffi module ``lib/pnp/_ext`` (built by ``lib/pnp/setup_ceres.py:17-31``).

The reference's ``lib/pnp/pnp_ceres.py:1-2`` does ``from ._ext import lib`` / ``from ._ext import ffi`` and calls
``lib.pnp_ceres_f32_omp`` (``pnp_ceres.py:136-139``) with the signature of ``lib/pnp/cxx/ext.h:1-14``.  ``liblc_b200.so``
exports that symbol (``lc_b200/csrc/lc_compat.cu``), so copying THIS file to ``lib/pnp/_ext.py`` (or aliasing it in
``sys.modules``) makes the reference's unmodified ``pnp_ceres.py`` / ``cer_solver.py`` run on the sm_100a solver; no
libceres, Eigen or glog needed.  cffi ABI mode: nothing is compiled.
"""
import cffi

from .. import _native

ffi = cffi.FFI()
# lib/pnp/cxx/ext.h:1-14 (the reference declares the return type void; the status code returned here may be ignored)
ffi.cdef("""
int pnp_ceres_f32_omp(
    float ** init_states,
    float ** cam_Ks,
    float ** pts2ds,
    float ** pts3ds,
    float ** icov_sqrtLs,
    int * ptCnts,
    int maxIterCnt,
    float function_tolerance,
    int printSummary,
    float* result_trs, int* rets,
    int job_count,
    int num_threads
);
""")
_native.lib()   # raises NativeLibraryError when liblc_b200.so has not been built (no fallback)
lib = ffi.dlopen(_native.LIB_PATH)
