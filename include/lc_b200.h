/*
 * lc_b200 — C ABI of the B200-native LC hot path (liblc_b200.so).
 *
 * This is the drop-in boundary: plain pointers, sizes, element strides and a CUDA stream;
 * no torch / C++ types.  Every pointer is a DEVICE pointer owned by the caller (the library
 * never allocates, frees or synchronises); kernels are enqueued on `stream` and the call
 * returns immediately.  Return value: 0 on success, otherwise the cudaError_t of the failed
 * launch / a negative LC_E_* code; lc_b200_last_error() gives the text.  Re-entrant.
 *
 * What each entry point replaces in the reference (fulliu/lc):
 *
 *   lc_b200_lm_solve        pnp_ceres_f32_omp / pnp_ceres_f32    lib/pnp/cxx/ext.h:2-15,
 *                           (cffi binding)                       lib/pnp/cxx/ceres.cpp:72-177,
 *                           + the host-side prologue of          lib/pnp/pnp_ceres.py:74-140,
 *                           cer_solver.solve (icov -> L,         lib/pnp/cer_solver.py:27-53
 *                           nan_to_num, invalid -> start)
 *   lc_b200_loss_fwd_bwd    Loss_cov_mixed forward + the         lib/cov_mixed.py:100-150
 *                           autograd backward it implies         (lib/nll/pnp_auto.py:86-135,
 *                                                                 lib/nll/pnp_utils.py:82-167)
 *   lc_b200_solve_loss      cer_solver.solve followed by         test.py:127 + losses.py:383
 *                           Loss_cov_mixed(pose := solution)     (fused, one launch)
 *   lc_b200_pnp_jac_cov     weighted_pnp_jac_wrt_pts2d(...,      lib/nll/pnp_auto.py:111-135
 *                           with_cov=True) forward
 *   lc_b200_pnp_jac_cov_bwd its double-backward w.r.t. weights   lib/nll/pnp_auto.py:129-134
 *   lc_b200_dense_loss_fwd_bwd  the producer glue of Loss_fn.dense_pose_loss fused in front of the loss
 *                           (gdr-net and zebrapose branches)     losses.py:336-386, 142-184, floatbits.py:49-160
 *   lc_b200_noc_bin_decode  nn_out_to_xyz(..., inference=True)   losses.py:16-45, floatbits.py:33-47, 197-224
 *
 * Layout: every array is described by an lc_view = base pointer + strides IN ELEMENTS, so the
 * planar (B,N,C) views the dense call site produces (strides (C*N,1,N), losses.py:142-161),
 * batch-broadcast grids (stride 0) and contiguous AoS tensors are all consumed without a copy.
 * `dtype` selects the element type of all floating arrays (LC_F32 or LC_F64); arithmetic is fp64.
 */
#ifndef LC_B200_H_
#define LC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC_B200_ABI_VERSION 2

enum { LC_F32 = 0, LC_F64 = 1 };

/* error codes (negative; positive values are cudaError_t) */
enum { LC_OK = 0, LC_E_BADARG = -1, LC_E_DTYPE = -2, LC_E_NULL = -3 };

/* weight interpretation for the solver (cer_solver.py:37-40, test.py:54,95) */
enum {
    LC_W_ICOV_DIAG = 0, /* (B,N,2) inverse variances;      L = diag(sqrt(icov))            */
    LC_W_ICOV_FULL = 1, /* (B,N,2,2) inverse covariances;  L = lower Cholesky factor       */
    LC_W_INV_STD = 2,   /* (B,N,2) inverse std (loss-half input); icov = inv_std^2          */
    LC_W_SQRT_L = 3     /* (B,N,2,2) L itself, row-major, element [0][1] ignored (ext.h ABI) */
};

/* flag bits */
enum {
    LC_FLAG_NAN_TO_NUM = 1,        /* solver: torch.nan_to_num on K, pts3d, pts2d, weights, start (cer_solver.py:27-29) */
    LC_FLAG_TOL_NEEDS_SUCCESS = 2, /* solver: Ceres >= 2.1 "atleast_one_successful_step" guard (see oracle/lm_oracle.c) */
    LC_FLAG_EXACT_HESSIAN = 4,     /* pnp_jac_cov(_bwd): add the r * d2r term of hessian_6d_elem (pnp_auto.py:59-83); needs pts2d */
    LC_FLAG_FORCE_STREAMING = 8,   /* per-pose entry points: always use the streaming fp64 kernel (tests / A-B comparisons)  */
    LC_FLAG_LM_MIXED = 16,         /* solver, resident kernels with planar fp32 weights: residuals, cost and every trust-region decision
                                      in fp64 as always, but the Jacobian rows and their sums J^T J, J^T r in packed fp32 per thread (fp64
                                      across threads).  Moves ~3/4 of the pass off the fp64 pipe.  The iterates differ from the all-fp64
                                      pass by ~1e-6 of a step (<< the 1e-6 rad tolerance); profiles/lm_mixed_check_r2.md holds the
                                      iteration-count / accept-flag comparison over 10k poses.  Opt-in; ignored by the other kernels. */
    LC_FLAG_COV_2D = 32            /* loss: Loss_cov_mixed(..., cov_2d=True) (lib/cov_mixed.py:76-80, 91-97): the corner covariances and the
                                      linear term are those of the PROJECTED bbox corners (8 x 2 rows through project_apply) instead of
                                      the 3-D ones (8 x 3).  No reference config enables it; served by the streaming kernel only
                                      (lc_b200_loss_fwd_bwd; the fused / dense entry points reject it). */
};

/* per-pose status bits written to `lc_flags` */
enum { LC_ST_HESS_NOT_SPD = 1, LC_ST_PRIOR_NOT_GOOD = 2, LC_ST_COV_NOT_GOOD = 4 };

typedef struct lc_view {
    void* ptr;         /* device pointer, may be NULL for optional arrays */
    int64_t stride[4]; /* element strides: [batch, point-or-row, component-or-col, (col)] */
} lc_view;

typedef struct lc_args {
    int32_t abi_version; /* LC_B200_ABI_VERSION */
    int32_t B;           /* poses */
    int32_t N;           /* correspondences per pose (padded maximum for ragged batches) */
    int32_t dtype;       /* LC_F32 | LC_F64 */
    int32_t flags;       /* LC_FLAG_* */
    int32_t weight_mode; /* LC_W_* (solver) */
    int32_t max_iter;    /* solver: max_num_iterations (50) */
    int32_t reserved0;
    double function_tolerance; /* solver (1e-6) */
    double max_err_len;        /* loss: clamp_error (32) */
    double rel_thresh;         /* loss: robust_weights_cov (3) */
    double w_e_thresh;         /* loss: robust_weights_cov (4) */
    double grad_scale;         /* loss: all input gradients are multiplied by grad_scale * grad_out[b] */

    /* inputs */
    lc_view K;        /* (B,3,3) */
    lc_view pose;     /* (B,7) wxyz+t: loss operating point / solver start */
    lc_view pts3d;    /* (B,N,3) */
    lc_view pts2d;    /* (B,N,2) */
    lc_view weights;  /* loss: inv_std (B,N,2); solver: see weight_mode; jac: weights (B,N,2) */
    lc_view valid;    /* (B,N) or NULL */
    lc_view bbox;     /* (B,8,3) */
    lc_view grad_out; /* (B) upstream d/d loss_b, or NULL (= 1) */
    const int32_t* n_points; /* (B) valid correspondences per pose, or NULL (= N); the gradient slots i >= n_points[b] are written as 0 */

    /* outputs (NULL = not wanted) */
    lc_view loss;      /* (B) */
    lc_view g_pts3d;   /* (B,N,3) */
    lc_view g_pts2d;   /* (B,N,2) */
    lc_view g_weights; /* (B,N,2) d/d inv_std (loss) or d/d weights (jac bwd) */
    lc_view cov;        /* (B,6,6) prior_update_cov = H^-1 */
    lc_view update_cov; /* (B,6,6) sym(A diag(sigma) A^T) */
    lc_view jac;        /* (B,6,N,2) d(update)/d(pts2d) */
    lc_view g_jac;      /* (B,6,N,2) upstream gradient of jac (jac bwd) */
    lc_view g_cov;      /* (B,6,6) upstream gradient of cov (jac bwd) or NULL */
    lc_view state;      /* (B,7) solver result (start where invalid) */
    lc_view radius;     /* (B) final trust-region radius */
    int32_t* invalid;   /* (B) solver_invalids */
    int32_t* iters;     /* (B) LM iterations taken, or NULL */
    int32_t* lc_flags;  /* (B) LC_ST_* bits, or NULL */
    double* trace;      /* (B, max_iter+2, 4) [cost, radius, step_ok, gmax] per finalised iteration, or NULL */
    double* loss_sum;   /* (2) += [sum_b loss_b, B] by atomicAdd (caller zeroes): the operand of the one scalar
                           all-reduce a batch-sharded mean needs (losses.py:334,386), or NULL */
} lc_args;

/*
 * Dense producer fused with the LC loss ("next" row f1 of SURVEY.md §8): replaces, for the gdr-net structure, the glue of
 * Loss_fn.dense_pose_loss (losses.py:336-386) in front of Loss_cov_mixed and its autograd backward:
 *   weights  = softmax(all 2*H*W weight logits of a sample) * weights_scale            (losses.py:355-356)
 *   inv_std  = weights[..., top::s, left::s]   pts3d = xyz_noc[..., top::s, left::s] * noc_scale
 *   pts2d    = gen_uv grid[top::s, left::s]    valid = ones                            (losses.py:142-161, 366)
 *   loss_b   = Loss_cov_mixed(K, pose, pts3d, pts2d, inv_std, valid, bbox_3d=..., max_err_len=...)   (losses.py:383)
 * One launch produces loss (B) and the gradients w.r.t. xyz_noc (B,3,H,W), the logits (B,2,H,W) and weights_scale (B)
 * (softmax backward included; un-sampled pixels get their exact gradient: 0 for xyz_noc, -w_j*S for the logits).
 * fp32 only.  The (H,W) planes of xyz_noc / logits / their gradients must be contiguous (stride[3]==1, stride[2]==W).
 */
typedef struct lc_dense_args {
    int32_t abi_version, B, H, W;
    int32_t sample, top, left, reserved0;
    double max_err_len, rel_thresh, w_e_thresh, grad_scale;
    lc_view xyz_noc;       /* (B,3,H,W) */
    lc_view logits;        /* (B,2,H,W) xyz_weight_logits */
    lc_view weights_scale; /* (B) xyz_weights_scale */
    lc_view noc_scale;     /* (B,3) */
    lc_view K, pose, bbox; /* (B,3,3), (B,7), (B,8,3) */
    lc_view grad_out;      /* (B) or NULL */
    lc_view loss;          /* (B) */
    lc_view g_xyz_noc;     /* (B,3,H,W) or NULL */
    lc_view g_logits;      /* (B,2,H,W) or NULL */
    lc_view g_scale;       /* (B) or NULL */
    lc_view cov, update_cov; /* (B,6,6) or NULL */
    int32_t* lc_flags;     /* (B) or NULL */
    double* loss_sum;      /* (2) or NULL, see lc_args.loss_sum */

    /*
     * ZebraPose producer ("next" row f3).  When noc_bin_logits.ptr != NULL, pts3d comes from the Gray-coded bit logits
     * instead of xyz_noc (which is then ignored, as is g_xyz_noc): the zebrapose branch of dense_pose_loss,
     *   dense_pnp_matching_from_noc_bin                      losses.py:163-184
     *   nn_out_to_xyz(raw_bits_gt=..., noc_mask=..., ...)   losses.py:16-45   (noc * noc_scale, model transform)
     *   floatbits.nn_logits2noc_with_gt                      floatbits.py:49-69, 99-160 (MSB-error soft decoding)
     * and its backward: per sampled in-mask pixel and axis exactly one bit channel receives a gradient.
     */
    lc_view noc_bin_logits;  /* (B,C,H,W) fp32, C = bit_cnt[0]+bit_cnt[1]+bit_cnt[2], contiguous (H,W) planes */
    lc_view noc_bin_raw;     /* (B,C,H,W) uint8 / bool ground-truth raw bits, any strides (nn_noc2target returns channel-last storage) */
    lc_view msk_noc;         /* (B,H,W) uint8 / bool, any strides */
    lc_view model_transform; /* (B,4,4) fp32 or NULL: xyz = (xyz_xformed - T[:3,3]) @ T[:3,:3] */
    lc_view g_noc_bin;       /* (B,C,H,W) fp32 or NULL, contiguous (H,W) planes */
    int32_t bit_cnt[3];      /* bits per axis (floatbits.calc_bit_count), 1..16 each */
    int32_t black_background; /* floatbits._black_background: the two leading bits of every axis are stored inverted */
} lc_dense_args;

int lc_b200_dense_loss_fwd_bwd(const lc_dense_args* a, void* cuda_stream);

/*
 * Test-time ZebraPose decode (row f3): nn_out_to_xyz(nn_out, noc_scale, model_transform=..., bit_cnt=..., inference=True)
 * (losses.py:16-45) = floatbits.nn_logits2noc without LUT (floatbits.py:33-47, 197-224: hard Gray bits -> binary, soft LSB),
 * times noc_scale, model transform.  One pass: reads C*4 bytes and writes 12 bytes per pixel.
 */
typedef struct lc_decode_args {
    int32_t abi_version, B, H, W;
    int32_t bit_cnt[3];
    int32_t black_background;
    lc_view noc_bin_logits;  /* (B,C,H,W) fp32, contiguous (H,W) planes */
    lc_view noc_scale;       /* (B,3) */
    lc_view model_transform; /* (B,4,4) or NULL */
    lc_view xyz;             /* out (B,H,W,3) fp32, any strides: [batch, row, col, component] */
} lc_decode_args;

int lc_b200_noc_bin_decode(const lc_decode_args* a, void* cuda_stream);

/*
 * ZebraPose target coding (row f3): floatbits.nn_noc2target (floatbits.py:13-31, 76-97), the producer of the ground-truth bits
 * the training decode consumes (losses.py:64, 130-136): per axis ints = round(clamp((noc+1)*max/2, 0, max)), MSB-first binary
 * `raw`, Gray code `mod` (mod_j = raw_j xor raw_{j-1}) with the two leading bits inverted under a black background.
 * Outputs are written channel-last, (B,H,W,C) bytes, the storage the reference's permuted views have.
 */
typedef struct lc_encode_args {
    int32_t abi_version, B, H, W;
    int32_t bit_cnt[3];
    int32_t black_background;
    lc_view noc;       /* (B,H,W,3) fp32, any strides [batch,row,col,component] */
    uint8_t* mod_bits; /* out (B,H,W,C) 0/1 bytes, or NULL */
    uint8_t* raw_bits; /* out (B,H,W,C) 0/1 bytes, or NULL */
} lc_encode_args;

int lc_b200_noc_bin_encode(const lc_encode_args* a, void* cuda_stream);

/*
 * Test-time point selection ("next" row f2): the selection half of test.solve_pnp_dense (test.py:67-119) on the device.
 *   weights  = softmax(weight logits over 2*H*W, or per channel when weights_scale is (B,2)) * weights_scale   test.py:84-88
 *   sub-sample every `sample`-th pixel from (0,0): pts2d grid, inv_std, xyz, seg mask                           losses.py:142-161
 *   inv_cov  = inv_std^2                                                                                        test.py:95
 *   valid    = mask | quantile | quantile_in_mask (torch.quantile, linear interpolation, fp32)                  test.py:36-45, 97-104
 *   v.nonzero() -> ragged lists -> cer_solver._batch_tensors zero padding                                       test.py:106-119,
 *                                                                                                               cer_solver.py:67-87
 * Output: zero-padded (B,Nmax,.) correspondences in selection order + n_points (B): the inputs of lc_b200_lm_solve, with
 * no device->host synchronisation.  Samples with fewer than min_points (4) selected points are padded with pseudo-random
 * indices (test.py:108-113 draws them from np.random).  `weights` may carry the precomputed inv_std map instead of
 * logits + scale (then the selection is bit-exact w.r.t. the reference given the same map).
 */
enum { LC_SEL_MASK = 0, LC_SEL_QUANTILE = 1, LC_SEL_QUANTILE_IN_MASK = 2 };

typedef struct lc_select_args {
    int32_t abi_version, B, H, W;
    int32_t sample, mode, scale_dim, min_points; /* mode: LC_SEL_*; scale_dim 1 (joint softmax) or 2 (per channel); min_points 4 */
    int32_t Nmax, reserved0;                     /* Nmax = ceil(H/sample)*ceil(W/sample) */
    float quantile, one_minus_quantile;          /* cfg.quantile and (float)(1 - cfg.quantile) */
    float seg_thresh, reserved1;                 /* cfg.seg_thresh (0.5) */
    lc_view xyz;           /* (B,H,W,3) fp32, any strides [batch,row,col,component] (nn_out_to_xyz output or an NCHW permute) */
    lc_view noc_scale;     /* (B,3) multiplier or NULL */
    lc_view weights;       /* (B,2,H,W) precomputed inv_std, any strides, or NULL */
    lc_view logits;        /* (B,2,H,W) weight logits, contiguous (H,W) planes (when weights is NULL) */
    lc_view weights_scale; /* (B,scale_dim) */
    lc_view msk_logits;    /* (B,H,W) msk_vis_logits, strides [batch,row,col] */
    lc_view pts3d;         /* out (B,Nmax,3) */
    lc_view pts2d;         /* out (B,Nmax,2) */
    lc_view inv_cov;       /* out (B,Nmax,2) */
    int32_t* index;        /* out (B,Nmax) sampled-point index of every slot (-1 = padding), or NULL */
    int32_t* n_points;     /* out (B) */
} lc_select_args;

int lc_b200_dense_select(const lc_select_args* a, void* cuda_stream);

/*
 * Device-side pose initialiser ("next" row f2): the role of lib/pnp/cv2_solver.solve (cv2_solver.py:6-88,
 * cv2.solvePnPRansac(EPNP, 150 iterations) per sample on the host) in test.solve_pnp / solve_pnp_dense (test.py:60, 120):
 * a start pose for the weighted LM solve and an inlier set for the 'weighted_filtered' branch (test.py:131-134).
 * Weighted DLT reduced to a 4x4 eigenproblem + Cauchy IRLS on the pixel reprojection error; NOT OpenCV's RANSAC (see
 * lc_init.cu).  fp32 arrays, fp64 arithmetic.
 */
typedef struct lc_init_args {
    int32_t abi_version, B, N, irls_rounds; /* irls_rounds: robust re-solves after the first (3) */
    float reproj_thresh, reserved0;         /* pixels: cv2 reprojectionError (Cauchy scale and inlier threshold) */
    lc_view K, pts3d, pts2d;                /* (B,3,3), (B,N,3), (B,N,2) */
    lc_view weights;                        /* (B,N,2) inverse variances (base weights) or NULL (= 1) */
    lc_view reproj_thresh_b;                /* (B) per-sample threshold (cfg.rel_reproj_err, test.py:115-117) or NULL */
    const int32_t* n_points;                /* (B) or NULL */
    lc_view state;                          /* out (B,7) wxyz + t */
    int32_t* invalid;                       /* out (B) or NULL: fewer than 6 points / degenerate system */
    uint8_t* inlier;                        /* out (B,N) or NULL: reprojection error < threshold under the returned pose */
    int32_t* n_inliers;                     /* out (B) or NULL */
} lc_init_args;

int lc_b200_pnp_init(const lc_init_args* a, void* cuda_stream);

/*
 * Pose-error metrics ("next" row f4): compute_pose_errors (lib/utils/evaluate.py:333-339) = error6d.add / adi / re / te
 * (lib/utils/error6d.py:87-159) for a batch of poses, one CTA per pose (the reference: numpy + scipy cKDTree in a
 * multiprocessing.Pool(6), evaluate.py:193-210).  All arrays fp64 like the reference's numpy code.
 */
typedef struct lc_eval_args {
    int32_t abi_version, B, M, reserved0; /* M = model points per pose when pts_count is NULL */
    lc_view R_est, t_est, R_gt, t_gt;     /* (B,3,3), (B,3), (B,3,3), (B,3) fp64 */
    lc_view pts;                          /* (P,3) fp64 model points, strides [point, component] */
    const int64_t* pts_offset;            /* (B) first model point of pose b in pts, or NULL (= 0) */
    const int32_t* pts_count;             /* (B) model points of pose b, or NULL (= M) */
    lc_view add, adi, re, te;             /* out (B) fp64, each may be NULL */
} lc_eval_args;

int lc_b200_pose_errors(const lc_eval_args* a, void* cuda_stream);

/*
 * Symmetric pose-candidate selection ("next" row f4): symmetry.select_pose_2d (mode 0, symmetry.py:8-31) and
 * select_pose_3d (mode 1, symmetry.py:33-56): per sample the candidate (of Kc) with the smallest mean error; fp32.
 */
typedef struct lc_candi_args {
    int32_t abi_version, B, N, Kc;
    int32_t mode, reserved0;   /* 0: 2-D reprojection error, 1: 3-D back-projection error */
    lc_view K;                 /* (B,3,3) */
    lc_view pts_a;             /* mode 0: pts3d (B,N,3); mode 1: pts3d_out (B,N,3) */
    lc_view pts_b;             /* mode 0: pts2d (B,N,2); mode 1: homo_z (B,N,3) */
    lc_view candi;             /* (B,Kc,3,4) */
    lc_view best;              /* out (B,3,4) or NULL */
    lc_view err;               /* out (B,Kc) mean error per candidate, or NULL */
    int32_t* best_index;       /* out (B) or NULL */
} lc_candi_args;

int lc_b200_select_pose(const lc_candi_args* a, void* cuda_stream);

int lc_b200_abi_version(void);
const char* lc_b200_last_error(void);

int lc_b200_lm_solve(const lc_args* a, void* cuda_stream);
int lc_b200_loss_fwd_bwd(const lc_args* a, void* cuda_stream);
int lc_b200_solve_loss(const lc_args* a, void* cuda_stream);
int lc_b200_pnp_jac_cov(const lc_args* a, void* cuda_stream);
int lc_b200_pnp_jac_cov_bwd(const lc_args* a, void* cuda_stream);

/* number of kernels the last call on this thread launched (bench.py's gpu_launches claim) */
int lc_b200_last_launch_count(void);
/* names of those kernels (template instantiations as dispatched), '+'-separated; valid until the next call on this thread */
const char* lc_b200_last_kernels(void);

/*
 * Binary-compatible replacement of the symbol the reference's cffi module binds (lib/pnp/cxx/ext.h:1-14, implemented in
 * lib/pnp/cxx/ceres.cpp:147-177 and called from lib/pnp/pnp_ceres.py:136-139): same name, same argument list, HOST pointer
 * tables with one ragged job per entry (float[7] wxyz+t in/out, float[9] K, float[2n], float[3n], float[4n] row-major 2x2
 * sqrt-information factors with element [1] ignored, ceres.cpp:17-28).  This is the one entry point that owns memory: it
 * packs the jobs into a cached pinned buffer (num_threads host threads), copies them to the current device, runs ONE
 * lc_b200_lm_solve launch and copies states / radii / flags back before returning (synchronous, like the reference).
 * ceres.cpp semantics kept: ptCnt < 3 -> rets = 1, result_trs = 1, state untouched (:84-91); rets / result_trs always
 * written (:134-136); states written back only when valid (:137-144).  printSummary prints one line per job.
 * The reference declares the function void; the int returned here (0 / cudaError_t / LC_E_*) can be ignored.
 */
int pnp_ceres_f32_omp(float** init_states, float** cam_Ks, float** pts2ds, float** pts3ds, float** icov_sqrtLs, int* ptCnts,
                      int maxIterCnt, float function_tolerance, int printSummary, float* result_trs, int* rets, int job_count,
                      int num_threads);
/* single-problem form, ceres.cpp:72-83 */
int pnp_ceres_f32(float* io_state_quat, const float* cam_K, const float* pts2d, const float* pts3d, const float* icov_sqrtL,
                  int ptCnt, int maxIterCnt, float function_tolerance, int printSummary, float* result_tr, int* ret);
/* frees the staging buffers cached by the two functions above */
void lc_b200_compat_release(void);

#ifdef __cplusplus
}
#endif
#endif /* LC_B200_H_ */
