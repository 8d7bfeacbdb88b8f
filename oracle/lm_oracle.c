/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.            *** PARITY UNPINNED ***
 *
 * CPU restatement (plain C, fp64) of the reference's weighted-PnP solver half:
 *   pnp_ceres_f32 / pnp_ceres_f32_omp     /root/reference/lib/pnp/cxx/ceres.cpp:72-145, 147-177
 *   ReprojectionError                      ceres.cpp:15-65
 * and of the third-party algorithm it delegates to, which is NOT in /root/reference:
 *   Ceres Solver 2.1.0 (pinned by scripts/build-ceres.sh:19-22), call sites ceres.cpp:37
 *   (AngleAxisRotatePoint), :59 (AutoDiffCostFunction<.,2,6>), :96 (QuaternionToAngleAxis),
 *   :104-130 (Problem/Options/Solve: DENSE_QR, max_num_iterations, function_tolerance),
 *   :131 (AngleAxisToQuaternion), :134-136 (Summary).
 *
 * libceres / Eigen / glog are absent from this image and there is no network, so the
 * reference extension cannot be built and the reference holds no test or golden vector for
 * this boundary.  What follows restates the published algorithm of
 * Ceres 2.1.0's TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseQRSolver and
 * ceres/rotation.h from memory; it is the working definition of "Ceres-faithful" for this
 * repo.  PARITY WITH libceres ITSELF IS UNPINNED — every report built on this file must say so.
 * What *is* checked (tests/test_oracle_cpu.py): forward-mode Jacobians against finite
 * differences, noise-free known answers, stationarity of the returned pose and agreement
 * with scipy.optimize.least_squares run to full convergence within the early-stop gap.
 *
 * Faithful details that matter for the answer (a 1e-6 function-tolerance stop leaves the
 * pose 1e-6..1e-4 rad from the optimum, so the schedule IS the result):
 *   - parameters x = [angle-axis(3), t(3)], global, no manifold; residuals via Jet<double,6>
 *   - Jacobi scaling 1/(1+||J_col||) computed once at iteration 0, applied to every J
 *   - LM diagonal = clamp(colnorm^2(J_scaled), 1e-6, 1e32), reused after a rejected step
 *   - step = -argmin || [J; sqrt(diag/radius)] y - [r; 0] ||  by Householder QR
 *   - model_cost_change = -(Js)'(r + Js/2); <= 0 -> invalid step: radius *= 0.5 (5 in a row = FAILURE)
 *   - parameter- then function-tolerance tests BEFORE acceptance, candidate discarded;
 *     both only once at least one step has succeeded (flag LM_TOL_NEEDS_SUCCESS, see below)
 *   - accept iff rho > 1e-3; radius /= max(1/3, 1-(2rho-1)^3) capped at 1e16; reject: radius /= dec, dec *= 2
 *   - gradient test |x - (x + (-g))|_inf <= 1e-10 after a successful step; radius <= 1e-32 -> CONVERGENCE
 *   - wrapper: ptCnt<3 -> invalid, tr=1; invalid = termination != CONVERGENCE; state written
 *     back (fp32) only when valid; reported radius = radius at the last finalised iteration.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Every rule of the trust-region loop that is restated from memory (not from ceres.cpp, not from the documented
 * Solver::Options defaults) is switchable, so that tools/lm_sensitivity.py can measure how much the returned pose
 * depends on it (profiles/lm_unpinned_sensitivity.md).  flags = LM_TOL_NEEDS_SUCCESS (1) is the working definition.
 *
 *   LM_TOL_NEEDS_SUCCESS      parameter/function-tolerance convergence only after at least one successful step
 *                             ("atleast_one_successful_step"); off = tested unconditionally
 *   LM_INVALID_DIV_DEC        invalid step (model_cost_change <= 0 / non-finite solve): radius /= dec, dec *= 2 like a
 *                             rejected step; default = LevenbergMarquardtStrategy::StepIsInvalid, radius *= 0.5
 *   LM_GRAD_TEST_ALWAYS       gradient-tolerance test after every finalised iteration; default = only after a
 *                             successful step
 *   LM_REPORT_CURRENT_RADIUS  reported trust-region radius = the strategy's radius at termination; default = the radius
 *                             recorded in the last finalised IterationSummary (iterations.back())
 *   LM_SOLVE_NORMAL_EQ        step from a Cholesky factorisation of J^T J + D^2 (what the CUDA kernel does); default =
 *                             Householder QR of [J; D] (DENSE_QR)
 *   LM_TOL_KEEP_CANDIDATE     a candidate that triggers the parameter/function tolerance is kept when it lowers the cost;
 *                             default = discarded (Minimize() returns before HandleSuccessfulStep)
 *   LM_GRAD_NORM_PLAIN        gradient max-norm = |g|_inf; default = |x - Plus(x, -g)|_inf evaluated in floating point
 *   LM_SCALE_NO_PLUS_ONE      Jacobi scaling 1/||J_col||; default = 1/(1 + ||J_col||)
 */
#define LM_TOL_NEEDS_SUCCESS 1
#define LM_INVALID_DIV_DEC 2
#define LM_GRAD_TEST_ALWAYS 4
#define LM_REPORT_CURRENT_RADIUS 8
#define LM_SOLVE_NORMAL_EQ 16
#define LM_TOL_KEEP_CANDIDATE 32
#define LM_GRAD_NORM_PLAIN 64
#define LM_SCALE_NO_PLUS_ONE 128

/* ---------------- Jet<double,6> ---------------- */
typedef struct { double v; double d[6]; } jet;
static inline jet jc(double v) { jet r; r.v = v; memset(r.d, 0, sizeof(r.d)); return r; }
static inline jet jvar(double v, int k) { jet r = jc(v); r.d[k] = 1.0; return r; }
static inline jet jadd(jet a, jet b) { jet r; r.v = a.v + b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
static inline jet jsub(jet a, jet b) { jet r; r.v = a.v - b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
static inline jet jmul(jet a, jet b) { jet r; r.v = a.v * b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.v * b.d[k] + a.d[k] * b.v; return r; }
static inline jet jscale(jet a, double s) { jet r; r.v = a.v * s; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] * s; return r; }
static inline jet jdiv(jet a, jet b) {
    /* ceres/jet.h: g/h = g * (1/h),  d = (dg - (g/h) dh) / h */
    jet r; const double ih = 1.0 / b.v; r.v = a.v * ih;
    for (int k = 0; k < 6; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ih;
    return r;
}
static inline jet jsqrt(jet a) { jet r; r.v = sqrt(a.v); const double s = 1.0 / (2.0 * r.v); for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] * s; return r; }
static inline jet jsin(jet a) { jet r; r.v = sin(a.v); const double c = cos(a.v); for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] * c; return r; }
static inline jet jcos(jet a) { jet r; r.v = cos(a.v); const double s = -sin(a.v); for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] * s; return r; }

/* ceres/rotation.h AngleAxisRotatePoint, on Jets */
static void aa_rotate_jet(const jet w[3], const double pt[3], jet out[3]) {
    const jet th2 = jadd(jadd(jmul(w[0], w[0]), jmul(w[1], w[1])), jmul(w[2], w[2]));
    if (th2.v > DBL_EPSILON) {
        const jet th = jsqrt(th2), ct = jcos(th), st = jsin(th), ith = jdiv(jc(1.0), th);
        const jet a[3] = {jmul(w[0], ith), jmul(w[1], ith), jmul(w[2], ith)};
        const jet axp[3] = {jsub(jscale(a[1], pt[2]), jscale(a[2], pt[1])),
                            jsub(jscale(a[2], pt[0]), jscale(a[0], pt[2])),
                            jsub(jscale(a[0], pt[1]), jscale(a[1], pt[0]))};
        const jet tmp = jmul(jadd(jadd(jscale(a[0], pt[0]), jscale(a[1], pt[1])), jscale(a[2], pt[2])), jsub(jc(1.0), ct));
        for (int i = 0; i < 3; ++i) out[i] = jadd(jadd(jscale(ct, pt[i]), jmul(axp[i], st)), jmul(a[i], tmp));
    } else {
        const jet axp[3] = {jsub(jscale(w[1], pt[2]), jscale(w[2], pt[1])),
                            jsub(jscale(w[2], pt[0]), jscale(w[0], pt[2])),
                            jsub(jscale(w[0], pt[1]), jscale(w[1], pt[0]))};
        for (int i = 0; i < 3; ++i) out[i] = jadd(jc(pt[i]), axp[i]);
    }
}

/* same function on plain doubles (what Ceres runs when no Jacobian is requested) */
static void aa_rotate(const double w[3], const double pt[3], double out[3]) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (th2 > DBL_EPSILON) {
        const double th = sqrt(th2), ct = cos(th), st = sin(th), ith = 1.0 / th;
        const double a[3] = {w[0] * ith, w[1] * ith, w[2] * ith};
        const double axp[3] = {a[1] * pt[2] - a[2] * pt[1], a[2] * pt[0] - a[0] * pt[2], a[0] * pt[1] - a[1] * pt[0]};
        const double tmp = (a[0] * pt[0] + a[1] * pt[1] + a[2] * pt[2]) * (1.0 - ct);
        for (int i = 0; i < 3; ++i) out[i] = pt[i] * ct + axp[i] * st + a[i] * tmp;
    } else {
        const double axp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
        for (int i = 0; i < 3; ++i) out[i] = pt[i] + axp[i];
    }
}

/* the captured per-observation constants of ReprojectionError (ceres.cpp:17-28) */
typedef struct { double u, v, a, b, c, X[3]; } obs_t;

typedef struct {
    int N;
    const obs_t* obs;
    double cam[6];
} problem_t;

/* Evaluate: cost = sum_blocks 0.5*|r_block|^2; optional residuals r[2N], J[2N*6] (row-major), g[6] = J^T r.
 * Returns 0 when a residual / Jacobian entry is not finite (Ceres: evaluation failure). */
static int evaluate(const problem_t* p, const double x[6], double* cost, double* r, double* J, double* g) {
    double c = 0.0;
    if (g) memset(g, 0, 6 * sizeof(double));
    if (!J) {
        for (int i = 0; i < p->N; ++i) {
            const obs_t* o = &p->obs[i];
            double q[3];
            aa_rotate(x, o->X, q);
            q[0] += x[3]; q[1] += x[4]; q[2] += x[5];
            const double up = (q[0] * p->cam[0] + q[1] * p->cam[1]) / q[2];
            const double vp = (q[0] * p->cam[3] + q[1] * p->cam[4]) / q[2];
            const double du = up - o->u, dv = vp - o->v;
            const double r0 = du * o->a + dv * o->b, r1 = dv * o->c;
            if (!isfinite(r0) || !isfinite(r1)) return 0;
            if (r) { r[2 * i] = r0; r[2 * i + 1] = r1; }
            c += 0.5 * (r0 * r0 + r1 * r1);
        }
        *cost = c;
        return 1;
    }
    jet w[3], tt[3];
    for (int k = 0; k < 3; ++k) { w[k] = jvar(x[k], k); tt[k] = jvar(x[3 + k], 3 + k); }
    for (int i = 0; i < p->N; ++i) {
        const obs_t* o = &p->obs[i];
        jet q[3];
        aa_rotate_jet(w, o->X, q);
        for (int k = 0; k < 3; ++k) q[k] = jadd(q[k], tt[k]);
        const jet up = jdiv(jadd(jscale(q[0], p->cam[0]), jscale(q[1], p->cam[1])), q[2]);
        const jet vp = jdiv(jadd(jscale(q[0], p->cam[3]), jscale(q[1], p->cam[4])), q[2]);
        const jet du = jsub(up, jc(o->u)), dv = jsub(vp, jc(o->v));
        const jet r0 = jadd(jscale(du, o->a), jscale(dv, o->b)), r1 = jscale(dv, o->c);
        if (!isfinite(r0.v) || !isfinite(r1.v)) return 0;
        for (int k = 0; k < 6; ++k)
            if (!isfinite(r0.d[k]) || !isfinite(r1.d[k])) return 0;
        r[2 * i] = r0.v; r[2 * i + 1] = r1.v;
        memcpy(J + (size_t)(2 * i) * 6, r0.d, 6 * sizeof(double));
        memcpy(J + (size_t)(2 * i + 1) * 6, r1.d, 6 * sizeof(double));
        c += 0.5 * (r0.v * r0.v + r1.v * r1.v);
        if (g) for (int k = 0; k < 6; ++k) g[k] += r0.d[k] * r0.v + r1.d[k] * r1.v;
    }
    *cost = c;
    return 1;
}

/* Householder QR least squares: min || A y - b ||, A is m x 6 row-major (destroyed), b length m (destroyed).
 * Returns 0 if the result is not finite. (DenseQRSolver; LAPACK dgeqr2-style reflectors.) */
static int qr_solve6(double* A, double* b, int m, double y[6]) {
    const int n = 6;
    for (int k = 0; k < n; ++k) {
        double nrm = 0;
        for (int i = k; i < m; ++i) nrm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        nrm = sqrt(nrm);
        if (nrm == 0.0) continue;
        const double alpha = A[(size_t)k * n + k];
        const double beta = alpha >= 0 ? -nrm : nrm;
        const double v0 = alpha - beta;
        /* v = [v0, A[k+1:,k]]; H = I - tau v v^T / (v0^2) with tau=(beta-alpha)/beta normalised form */
        const double tau = (beta - alpha) / beta;
        const double inv_v0 = 1.0 / v0;
        for (int i = k + 1; i < m; ++i) A[(size_t)i * n + k] *= inv_v0;   /* store v (v_k = 1) */
        A[(size_t)k * n + k] = beta;
        for (int j = k + 1; j < n; ++j) {
            double s = A[(size_t)k * n + j];
            for (int i = k + 1; i < m; ++i) s += A[(size_t)i * n + k] * A[(size_t)i * n + j];
            s *= tau;
            A[(size_t)k * n + j] -= s;
            for (int i = k + 1; i < m; ++i) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k];
        }
        double s = b[k];
        for (int i = k + 1; i < m; ++i) s += A[(size_t)i * n + k] * b[i];
        s *= tau;
        b[k] -= s;
        for (int i = k + 1; i < m; ++i) b[i] -= s * A[(size_t)i * n + k];
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int j = k + 1; j < n; ++j) s -= A[(size_t)k * n + j] * y[j];
        y[k] = s / A[(size_t)k * n + k];
    }
    for (int k = 0; k < n; ++k) if (!isfinite(y[k])) return 0;
    return 1;
}

/* LM_SOLVE_NORMAL_EQ: (J^T J + diag(d2)) y = J^T r by a 6x6 Cholesky factorisation.  Returns 0 on breakdown. */
static int chol_solve6(const double* J, const double* r, int m, const double d2[6], double y[6]) {
    double A[36], g[6], L[36];
    memset(A, 0, sizeof(A)); memset(g, 0, sizeof(g)); memset(L, 0, sizeof(L));
    for (int i = 0; i < m; ++i)
        for (int a = 0; a < 6; ++a) {
            g[a] += J[(size_t)i * 6 + a] * r[i];
            for (int b = a; b < 6; ++b) A[a * 6 + b] += J[(size_t)i * 6 + a] * J[(size_t)i * 6 + b];
        }
    for (int a = 0; a < 6; ++a) { A[a * 6 + a] += d2[a]; for (int b = 0; b < a; ++b) A[a * 6 + b] = A[b * 6 + a]; }
    for (int j = 0; j < 6; ++j) {
        double d = A[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0.0) || !isfinite(d)) return 0;
        L[j * 6 + j] = sqrt(d);
        for (int i = j + 1; i < 6; ++i) {
            double v = A[i * 6 + j];
            for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = v / L[j * 6 + j];
        }
    }
    double z[6];
    for (int i = 0; i < 6; ++i) { double v = g[i]; for (int k = 0; k < i; ++k) v -= L[i * 6 + k] * z[k]; z[i] = v / L[i * 6 + i]; }
    for (int i = 5; i >= 0; --i) { double v = z[i]; for (int k = i + 1; k < 6; ++k) v -= L[k * 6 + i] * y[k]; y[i] = v / L[i * 6 + i]; }
    for (int k = 0; k < 6; ++k) if (!isfinite(y[k])) return 0;
    return 1;
}

/* ceres/rotation.h */
static void quat_to_aa(const double q[4], double aa[3]) {
    const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (s2 > 0.0) {
        const double s = sqrt(s2), c = q[0];
        const double two_theta = 2.0 * ((c < 0.0) ? atan2(-s, -c) : atan2(s, c));
        const double k = two_theta / s;
        aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
    } else {
        aa[0] = q[1] * 2.0; aa[1] = q[2] * 2.0; aa[2] = q[3] * 2.0;
    }
}
static void aa_to_quat(const double aa[3], double q[4]) {
    const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (th2 > 0.0) {
        const double th = sqrt(th2), h = th * 0.5, k = sin(h) / th;
        q[0] = cos(h); q[1] = aa[0] * k; q[2] = aa[1] * k; q[3] = aa[2] * k;
    } else {
        q[0] = 1.0; q[1] = aa[0] * 0.5; q[2] = aa[1] * 0.5; q[3] = aa[2] * 0.5;
    }
}

enum { TERM_CONVERGENCE = 0, TERM_NO_CONVERGENCE = 1, TERM_FAILURE = 2 };

static double grad_max_norm(const double x[6], const double g[6], int flags) {
    double gmax = 0;
    for (int k = 0; k < 6; ++k) {
        double d;
        if (flags & LM_GRAD_NORM_PLAIN) d = fabs(g[k]);
        else { const volatile double xs = x[k] + (-g[k]); d = fabs(x[k] - xs); }
        if (d > gmax) gmax = d;
    }
    return gmax;
}

/* Trace row per finalised iteration: [cost, radius, step_successful, gradient_max_norm] */
#define LM_TRACE_COLS 4

/*
 * One problem, same signature role as pnp_ceres_f32 (ceres.cpp:72-83) plus diagnostics.
 *   io_state7: wxyz + t (fp32, in/out)   K9: row-major 3x3 (only [0..5] used)   L4: N x (2x2 row-major), [1] ignored
 *   out_iters: number of finalised iterations after iteration 0;  out_term: TERM_*
 *   trace: optional (max_iter+2) x LM_TRACE_COLS;  work: >= (2N+6)*7 + 2N*6 + 2N*2 doubles + N obs_t
 */
void lm_oracle_pose(float* io_state7, const float* K9, const float* pts2d, const float* pts3d, const float* L4,
                    int N, int max_iter, float function_tolerance, int flags, float* out_radius, int* out_invalid,
                    int* out_iters, int* out_term, double* out_x6, double* trace) {
    if (out_iters) *out_iters = 0;
    if (out_term) *out_term = TERM_FAILURE;
    if (N < 3) { *out_invalid = 1; *out_radius = 1.0f; return; }

    double x[6];
    {
        const double q[4] = {io_state7[0], io_state7[1], io_state7[2], io_state7[3]};
        quat_to_aa(q, x);
        for (int i = 0; i < 3; ++i) x[3 + i] = io_state7[4 + i];
    }
    problem_t P;
    P.N = N;
    for (int i = 0; i < 6; ++i) P.cam[i] = K9[i];
    obs_t* obs = (obs_t*)malloc(sizeof(obs_t) * (size_t)N);
    for (int i = 0; i < N; ++i) {
        obs[i].u = (double)pts2d[2 * i] - P.cam[2];
        obs[i].v = (double)pts2d[2 * i + 1] - P.cam[5];
        obs[i].a = L4[4 * i]; obs[i].b = L4[4 * i + 2]; obs[i].c = L4[4 * i + 3];
        for (int k = 0; k < 3; ++k) obs[i].X[k] = pts3d[3 * i + k];
    }
    P.obs = obs;
    const int m = 2 * N;
    double* r = (double*)malloc(sizeof(double) * ((size_t)m + (size_t)m * 6 + (size_t)(m + 6) * 7));
    double* J = r + m;
    double* lhs = J + (size_t)m * 6;
    double* rhs = lhs + (size_t)(m + 6) * 6;

    const double ftol = (double)function_tolerance;
    const double gtol = 1e-10, ptol = 1e-8, min_rel_dec = 1e-3;
    const double min_radius = 1e-32, max_radius = 1e16, min_diag = 1e-6, max_diag = 1e32;
    double radius = 1e4, dec = 2.0;
    int reuse_diag = 0, n_invalid = 0, term = TERM_FAILURE, it = 0, step_ok, any_success = 0;
    double cost, g[6], scale[6], diag[6], gmax, xnorm;
    double best[6];
    memcpy(best, x, sizeof(best));
    double reported_radius = radius;
    int have_iter = 0;

    /* ---- iteration 0 ---- */
    if (!evaluate(&P, x, &cost, r, J, g)) goto done;
    for (int k = 0; k < 6; ++k) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + k] * J[(size_t)i * 6 + k];
        scale[k] = 1.0 / (((flags & LM_SCALE_NO_PLUS_ONE) ? 0.0 : 1.0) + sqrt(s));
    }
    for (int i = 0; i < m; ++i) for (int k = 0; k < 6; ++k) J[(size_t)i * 6 + k] *= scale[k];
    gmax = grad_max_norm(x, g, flags);
    xnorm = 0; for (int k = 0; k < 6; ++k) xnorm += x[k] * x[k]; xnorm = sqrt(xnorm);
    step_ok = 1;

    for (;;) {
        /* FinalizeIterationAndCheckIfMinimizerCanContinue */
        reported_radius = radius; have_iter = 1;
        if (step_ok) memcpy(best, x, sizeof(best));
        if (trace) { double* tr = trace + (size_t)it * LM_TRACE_COLS; tr[0] = cost; tr[1] = radius; tr[2] = step_ok; tr[3] = gmax; }
        if (it >= max_iter) { term = TERM_NO_CONVERGENCE; break; }
        if ((step_ok || (flags & LM_GRAD_TEST_ALWAYS)) && gmax <= gtol) { term = TERM_CONVERGENCE; break; }
        if (radius <= min_radius) { term = TERM_CONVERGENCE; break; }
        ++it;

        /* ComputeTrustRegionStep */
        if (!reuse_diag) {
            for (int k = 0; k < 6; ++k) {
                double s = 0;
                for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + k] * J[(size_t)i * 6 + k];
                diag[k] = fmin(fmax(s, min_diag), max_diag);
            }
        }
        double step[6];
        int valid;
        if (flags & LM_SOLVE_NORMAL_EQ) {
            double d2[6];
            for (int k = 0; k < 6; ++k) d2[k] = diag[k] / radius;
            valid = chol_solve6(J, r, m, d2, step);
        } else {
            memcpy(lhs, J, sizeof(double) * (size_t)m * 6);
            memset(lhs + (size_t)m * 6, 0, sizeof(double) * 36);
            for (int k = 0; k < 6; ++k) lhs[(size_t)(m + k) * 6 + k] = sqrt(diag[k] / radius);
            memcpy(rhs, r, sizeof(double) * m);
            memset(rhs + m, 0, sizeof(double) * 6);
            valid = qr_solve6(lhs, rhs, m + 6, step);
        }
        reuse_diag = 1;
        double model_change = 0;
        if (valid) {
            for (int k = 0; k < 6; ++k) step[k] = -step[k];
            for (int i = 0; i < m; ++i) {
                double js = 0;
                for (int k = 0; k < 6; ++k) js += J[(size_t)i * 6 + k] * step[k];
                model_change -= js * (r[i] + js / 2.0);
            }
            valid = model_change > 0.0;
        }
        if (!valid) {
            if (++n_invalid >= 5) { term = TERM_FAILURE; have_iter = 1; break; }
            if (flags & LM_INVALID_DIV_DEC) { radius = radius / dec; dec *= 2.0; }
            else radius *= 0.5;                              /* StepIsInvalid */
            reuse_diag = 1;
            step_ok = 0;
            continue;
        }
        n_invalid = 0;
        double delta[6], xc[6], cost_c;
        for (int k = 0; k < 6; ++k) { delta[k] = step[k] * scale[k]; xc[k] = x[k] + delta[k]; }
        if (!evaluate(&P, xc, &cost_c, NULL, NULL, NULL)) cost_c = DBL_MAX;

        const int tol_armed = !(flags & LM_TOL_NEEDS_SUCCESS) || any_success;
        /* ParameterToleranceReached */
        double sn = 0; for (int k = 0; k < 6; ++k) sn += (x[k] - xc[k]) * (x[k] - xc[k]); sn = sqrt(sn);
        if (tol_armed && (sn <= ptol * (xnorm + ptol) || fabs(cost - cost_c) <= ftol * cost)) {   /* parameter, then function tolerance */
            if ((flags & LM_TOL_KEEP_CANDIDATE) && cost_c < cost) memcpy(best, xc, sizeof(best));
            term = TERM_CONVERGENCE;
            break;
        }

        const double rho = cost_c >= DBL_MAX ? -DBL_MAX : (cost - cost_c) / model_change;
        if (rho > min_rel_dec) {
            memcpy(x, xc, sizeof(x));
            xnorm = 0; for (int k = 0; k < 6; ++k) xnorm += x[k] * x[k]; xnorm = sqrt(xnorm);
            if (!evaluate(&P, x, &cost, r, J, g)) { term = TERM_FAILURE; break; }
            for (int i = 0; i < m; ++i) for (int k = 0; k < 6; ++k) J[(size_t)i * 6 + k] *= scale[k];
            gmax = grad_max_norm(x, g, flags);
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3));
            radius = fmin(max_radius, radius);
            dec = 2.0; reuse_diag = 0; step_ok = 1; any_success = 1;
        } else {
            radius = radius / dec; dec *= 2.0; reuse_diag = 1; step_ok = 0;
        }
    }
done:
    (void)have_iter;
    if (out_iters) *out_iters = it;
    if (out_term) *out_term = term;
    if (out_x6) memcpy(out_x6, best, sizeof(best));
    *out_radius = (float)((flags & LM_REPORT_CURRENT_RADIUS) ? radius : reported_radius);
    *out_invalid = (term != TERM_CONVERGENCE);
    if (term == TERM_CONVERGENCE) {
        double q[4];
        aa_to_quat(best, q);
        for (int i = 0; i < 4; ++i) io_state7[i] = (float)q[i];
        for (int i = 0; i < 3; ++i) io_state7[4 + i] = (float)best[3 + i];
    }
    free(r);
    free(obs);
}

/* Batch driver over contiguous padded arrays (what cer_solver._batch_tensors builds, cer_solver.py:67-87),
 * one problem per OpenMP thread exactly like pnp_ceres_f32_omp (ceres.cpp:147-177). */
void lm_oracle_batch(int B, int Nmax, float* io_states, const float* Ks, const float* pts2d, const float* pts3d,
                     const float* L, const int* n_points, int max_iter, float function_tolerance, int flags,
                     float* out_radius, int* out_invalid, int* out_iters, int* out_term, double* out_x6,
                     double* trace /* B x (max_iter+2) x 4 or NULL */, int threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const size_t n = (size_t)Nmax;
        lm_oracle_pose(io_states + 7 * b, Ks + 9 * b, pts2d + 2 * n * b, pts3d + 3 * n * b, L + 4 * n * b,
                       n_points ? n_points[b] : Nmax, max_iter, function_tolerance, flags, out_radius + b,
                       out_invalid + b, out_iters ? out_iters + b : NULL, out_term ? out_term + b : NULL,
                       out_x6 ? out_x6 + 6 * b : NULL,
                       trace ? trace + (size_t)b * (max_iter + 2) * LM_TRACE_COLS : NULL);
    }
}

/* Exposed for the Jacobian finite-difference test: residuals + Jacobian at x for one problem. */
int lm_oracle_eval(const double x[6], const float* K9, const float* pts2d, const float* pts3d, const float* L4, int N,
                   double* cost, double* r, double* J, double* g) {
    problem_t P;
    P.N = N;
    for (int i = 0; i < 6; ++i) P.cam[i] = K9[i];
    obs_t* obs = (obs_t*)malloc(sizeof(obs_t) * (size_t)N);
    for (int i = 0; i < N; ++i) {
        obs[i].u = (double)pts2d[2 * i] - P.cam[2];
        obs[i].v = (double)pts2d[2 * i + 1] - P.cam[5];
        obs[i].a = L4[4 * i]; obs[i].b = L4[4 * i + 2]; obs[i].c = L4[4 * i + 3];
        for (int k = 0; k < 3; ++k) obs[i].X[k] = pts3d[3 * i + k];
    }
    P.obs = obs;
    const int ok = evaluate(&P, x, cost, r, J, g);
    free(obs);
    return ok;
}
