/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64) of the reference's LC-loss half:
 *   Loss_cov_mixed                     /root/reference/lib/cov_mixed.py:100-150
 *   clamp_error / twice_huber /        lib/cov_mixed.py:16-24, 10-13, 27-39
 *     robust_weights_cov
 *   weighted_pnp_jac_wrt_pts2d,        lib/nll/pnp_auto.py:111-135, 86-108, 13-56
 *     diff_pnp_perturb, residual_with_jac6d
 *   safe_cholesky (non-SPD -> I)       lib/nll/pnp_utils.py:140-167
 *   jac_update2alter, xform_3d,        lib/cov_mixed.py:52-65, 73-75, 68-70, 83-89
 *     transformed_cov_from_jac, loss_cov_3d
 *   quaternion_to_matrix (2/|q| quirk) lib/transforms/rotation_conversions.py:39-68
 *   project_apply (z clamp 0.1)        lib/transforms/transforms.py:47-63
 *
 * The reference evaluates this with functorch (vmap/jacfwd + ~30 autograd sweeps);
 * here it is the closed form of SURVEY.md §8a ("LC math"): three 6x6 / 6-vector
 * accumulators H, G, b, 6x6 algebra, and a manual reverse pass.  Pinned against the
 * reference itself: the tests/golden npz fixtures are produced by tests/golden/make_golden.py
 * running the unmodified reference in fp64; tests/test_oracle_cpu.py checks this
 * file against them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void quat_to_R_ref(const double q[4], double R[9]) {
    /* rotation_conversions.py:39-68 — note two_s = 2/|q| (not 2/|q|^2) */
    const double r = q[0], i = q[1], j = q[2], k = q[3];
    const double two_s = 2.0 / sqrt(r * r + i * i + j * j + k * k);
    R[0] = 1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r); R[4] = 1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = 1 - two_s * (i * i + j * j);
}

/* lower Cholesky of a 6x6 SPD matrix, returns LAPACK-style info (0 = ok) */
static int chol6(const double A[36], double L[36]) {
    memset(L, 0, 36 * sizeof(double));
    for (int j = 0; j < 6; ++j) {
        double d = A[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0.0) || !isfinite(d)) return j + 1;
        const double ljj = sqrt(d);
        L[j * 6 + j] = ljj;
        for (int i = j + 1; i < 6; ++i) {
            double v = A[i * 6 + j];
            for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = v / ljj;
        }
    }
    return 0;
}

/* C = (L L^T)^-1 */
static void chol6_inverse(const double L[36], double C[36]) {
    double Li[36];
    memset(Li, 0, sizeof(Li));
    for (int c = 0; c < 6; ++c) {           /* invert lower-triangular L column by column */
        Li[c * 6 + c] = 1.0 / L[c * 6 + c];
        for (int r = c + 1; r < 6; ++r) {
            double v = 0;
            for (int k = c; k < r; ++k) v -= L[r * 6 + k] * Li[k * 6 + c];
            Li[r * 6 + c] = v / L[r * 6 + r];
        }
    }
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
            double v = 0;
            for (int k = 0; k < 6; ++k) v += Li[k * 6 + a] * Li[k * 6 + b];
            C[a * 6 + b] = v;
        }
}

static void mm6(const double A[36], const double B[36], double O[36]) {
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
            double v = 0;
            for (int k = 0; k < 6; ++k) v += A[a * 6 + k] * B[k * 6 + b];
            O[a * 6 + b] = v;
        }
}

/* Jacobian rows of residual_with_jac6d at zero perturbation (pnp_auto.py:33-54):
 * J (2x6) = K[:2,:2] . (1/z)[I2 | -uv0] . [ R(-[X]x) | I3 ],  z = camera-frame depth, no clamp */
static void point_jac(const double K[9], const double R[9], const double t[3], const double X[3], double J[12]) {
    const double P0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double P1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double P2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double iz = 1.0 / P2, u0 = P0 * iz, v0 = P1 * iz;
    /* D = K2x2 . iz [I2 | -uv0]  (2x3) */
    double D[6];
    for (int a = 0; a < 2; ++a) {
        const double k0 = K[a * 3 + 0], k1 = K[a * 3 + 1];
        D[a * 3 + 0] = k0 * iz; D[a * 3 + 1] = k1 * iz; D[a * 3 + 2] = -(k0 * u0 + k1 * v0) * iz;
    }
    /* M = R(-[X]x): column j of -[X]x is -(X x e_j) = e_j x X */
    const double nX[9] = {0, X[2], -X[1], -X[2], 0, X[0], X[1], -X[0], 0};   /* -[X]x row-major */
    double M[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[r * 3 + c] = R[r * 3 + 0] * nX[0 * 3 + c] + R[r * 3 + 1] * nX[1 * 3 + c] + R[r * 3 + 2] * nX[2 * 3 + c];
    for (int a = 0; a < 2; ++a) {
        for (int c = 0; c < 3; ++c)
            J[a * 6 + c] = D[a * 3 + 0] * M[0 * 3 + c] + D[a * 3 + 1] * M[1 * 3 + c] + D[a * 3 + 2] * M[2 * 3 + c];
        for (int c = 0; c < 3; ++c) J[a * 6 + 3 + c] = D[a * 3 + c];
    }
}

/* rho(S) of SURVEY §8a: mean_j sqrt(good ? sum_xyz diag(Jb_j S Jb_j^T) : 1); also returns per-corner sums */
static double rho_cov(const double Jb[8][18], int nd, const double S[36], double sj[8], int* good) {
    *good = 1;
    for (int j = 0; j < 8; ++j) {
        sj[j] = 0;
        for (int r = 0; r < nd; ++r) {
            const double* row = &Jb[j][r * 6];
            double v = 0;
            for (int a = 0; a < 6; ++a) {
                double w = 0;
                for (int b = 0; b < 6; ++b) w += S[a * 6 + b] * row[b];
                v += row[a] * w;
            }
            if (!(v > 0.0)) *good = 0;
            sj[j] += v;
        }
    }
    double acc = 0;
    for (int j = 0; j < 8; ++j) acc += sqrt(*good ? sj[j] : 1.0);
    return acc / 8.0;
}

/*
 * One pose.  AoS fp64 inputs: X[N*3], x[N*2], s[N*2], valid[N] or NULL, bbox[24].
 * Outputs (any may be NULL): loss[1], gX[N*3], gx[N*2], gs[N*2] (for d loss = 1),
 * A[6*N*2] = jac_pts2update, C[36] = prior_update_cov, M[36] = update_cov,
 * Wout[N*2], sig_out[N*2], flag[1] (bit0: Hessian not SPD -> identity; bit1: !good^C; bit2: !good^M).
 */
void lc_oracle_pose_ex(const double* K, const double* pose, const double* X, const double* x, const double* s,
                       const double* valid, const double* bbox, int N, double Lmax, double rel, double we,
                       double* loss, double* gX, double* gx, double* gs, double* A, double* C_out, double* M_out,
                       double* Wout, double* sig_out, int* flag, double* scratch /* >= 8*N doubles */,
                       int cov2d /* cov_mixed.py:76-80: corner covariances of the projected bbox instead of the 3-D one */) {
    double R[9];
    quat_to_R_ref(pose, R);
    const double* t = pose + 4;
    double* ec = scratch;             /* N*2 */
    double* sig = scratch + 2 * N;    /* N*2 */
    double* W = scratch + 4 * N;      /* N*2 */
    double* proj = scratch + 6 * N;   /* N*2 */
    int fl = 0;

    /* pass 1: project_apply (transforms.py:58-63), clamp_error (cov_mixed.py:16-24), sum |ec| */
    double m[2] = {0, 0}, vcnt = 0;
    for (int i = 0; i < N; ++i) {
        const double* Xi = X + 3 * i;
        double P[3], KP[3];
        for (int r = 0; r < 3; ++r) P[r] = R[r * 3] * Xi[0] + R[r * 3 + 1] * Xi[1] + R[r * 3 + 2] * Xi[2] + t[r];
        for (int r = 0; r < 3; ++r) KP[r] = K[r * 3] * P[0] + K[r * 3 + 1] * P[1] + K[r * 3 + 2] * P[2];
        const double zc = KP[2] > 0.1 ? KP[2] : 0.1;
        proj[2 * i] = KP[0] / zc; proj[2 * i + 1] = KP[1] / zc;
        double e0 = x[2 * i] - proj[2 * i], e1 = x[2 * i + 1] - proj[2 * i + 1];
        const double len = sqrt(e0 * e0 + e1 * e1) + 1e-6;
        const double f = (len - Lmax) / len;
        if (f > 0) { e0 -= f * e0; e1 -= f * e1; }
        ec[2 * i] = e0; ec[2 * i + 1] = e1;
        const double vi = valid ? valid[i] : 1.0;
        m[0] += vi * fabs(e0); m[1] += vi * fabs(e1); vcnt += vi;
    }
    if (!valid) vcnt = (double)N;
    m[0] /= vcnt; m[1] /= vcnt;
    const double d[2] = {rel * m[0], rel * m[1]};

    /* pass 2: sigma = twice_huber(|ec|, d); q = mean(s^2 sigma)  (cov_mixed.py:32-36) */
    double q[2] = {0, 0};
    for (int i = 0; i < N; ++i) {
        const double vi = valid ? valid[i] : 1.0;
        for (int a = 0; a < 2; ++a) {
            const double av = fabs(ec[2 * i + a]);
            const double sg = av > d[a] ? d[a] * (2 * av - d[a]) : av * av;
            sig[2 * i + a] = sg;
            q[a] += vi * (s[2 * i + a] * s[2 * i + a]) * sg;
        }
    }
    q[0] /= vcnt; q[1] /= vcnt;

    /* pass 3: W = twice_huber(s, delta), accumulate H, G, b  (cov_mixed.py:37-38, pnp_auto.py:86-100) */
    double H[36], G[36], bvec[6];
    memset(H, 0, sizeof(H)); memset(G, 0, sizeof(G)); memset(bvec, 0, sizeof(bvec));
    for (int i = 0; i < N; ++i) {
        double J[12];
        point_jac(K, R, t, X + 3 * i, J);
        for (int a = 0; a < 2; ++a) {
            const int k = 2 * i + a;
            const double del = sqrt((q[a] * we) / (sig[k] + 1e-6));
            const double sk = s[k];
            const double w = sk > del ? del * (2 * sk - del) : sk * sk;
            W[k] = w;
            const double wg = w * w * sig[k], wb = w * ec[k];
            const double* Jk = J + 6 * a;
            for (int r = 0; r < 6; ++r) {
                for (int c = 0; c < 6; ++c) {
                    const double jj = Jk[r] * Jk[c];
                    H[r * 6 + c] += w * jj;
                    G[r * 6 + c] += wg * jj;
                }
                bvec[r] += wb * Jk[r];
            }
        }
    }
    if (Wout) memcpy(Wout, W, sizeof(double) * 2 * N);
    if (sig_out) memcpy(sig_out, sig, sizeof(double) * 2 * N);

    /* safe_cholesky: non-SPD -> identity (pnp_utils.py:140-167) */
    double L[36], Cm[36];
    if (chol6(H, L) != 0) {
        fl |= 1;
        memset(H, 0, sizeof(H));
        for (int a = 0; a < 6; ++a) H[a * 6 + a] = 1.0;
        chol6(H, L);
    }
    chol6_inverse(L, Cm);

    double T1[36], Mm[36], dth[6];
    mm6(Cm, G, T1); mm6(T1, Cm, Mm);
    for (int a = 0; a < 6; ++a) { double v = 0; for (int k = 0; k < 6; ++k) v += Cm[a * 6 + k] * bvec[k]; dth[a] = v; }
    if (C_out) memcpy(C_out, Cm, sizeof(Cm));
    if (M_out) memcpy(M_out, Mm, sizeof(Mm));

    if (A)
        for (int i = 0; i < N; ++i) {
            double J[12];
            point_jac(K, R, t, X + 3 * i, J);
            for (int a = 0; a < 2; ++a)
                for (int r = 0; r < 6; ++r) {
                    double v = 0;
                    for (int k = 0; k < 6; ++k) v += Cm[r * 6 + k] * J[a * 6 + k];
                    A[(size_t)r * 2 * N + 2 * i + a] = W[2 * i + a] * v;
                }
        }

    /* bbox corner Jacobians Jb_j = [R(-[c_j]x) | I3]  (cov_mixed.py:52-65, 73-75) */
    /* The reference differentiates quaternion_to_matrix(q (x) dq(delta)) with its 2/|q| scaling; for a
     * quaternion of norm n (fp32-rounded poses are never exactly unit) that derivative is
     * n * R_true (-[c]x) with R_true = I + (R_ref - I)/n, not R_ref (-[c]x).  Identical when n = 1. */
    double Jb[8][18], Rb[9];
    {
        const double nq = sqrt(pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2] + pose[3] * pose[3]);
        for (int k = 0; k < 9; ++k) { const double id = (k % 4 == 0) ? 1.0 : 0.0; Rb[k] = nq * id + (R[k] - id); }
    }
    for (int j = 0; j < 8; ++j) {
        const double* c = bbox + 3 * j;
        const double nC[9] = {0, c[2], -c[1], -c[2], 0, c[0], c[1], -c[0], 0};
        for (int r = 0; r < 3; ++r) {
            for (int cc = 0; cc < 3; ++cc)
                Jb[j][r * 6 + cc] = Rb[r * 3] * nC[cc] + Rb[r * 3 + 1] * nC[3 + cc] + Rb[r * 3 + 2] * nC[6 + cc];
            for (int cc = 0; cc < 3; ++cc) Jb[j][r * 6 + 3 + cc] = (r == cc) ? 1.0 : 0.0;
        }
    }
    /* cov_2d (cov_mixed.py:76-80, 91-97, 129-131): the rows are those of the PROJECTED corners, xform_2d = project_apply(K, R c + t):
     * row2d[j][a] = sum_k dproj_a/dP_k row3d[j][k],  dproj/dP = (K[:2,:] - proj (x) K[2,:] [z > 0.1]) / max(z, 0.1);  two per corner. */
    const int nd = cov2d ? 2 : 3;
    if (cov2d) {
        for (int j = 0; j < 8; ++j) {
            const double* c = bbox + 3 * j;
            double P[3], KP[3], r3[18];
            for (int r = 0; r < 3; ++r) P[r] = R[r * 3] * c[0] + R[r * 3 + 1] * c[1] + R[r * 3 + 2] * c[2] + t[r];
            for (int r = 0; r < 3; ++r) KP[r] = K[r * 3] * P[0] + K[r * 3 + 1] * P[1] + K[r * 3 + 2] * P[2];
            const int act = KP[2] > 0.1;   /* clamp(min=0.1): gradient passes where z >= min; equality has measure zero */
            const double zc = act ? KP[2] : 0.1;
            memcpy(r3, Jb[j], sizeof(r3));
            for (int a2 = 0; a2 < 2; ++a2) {
                const double pr = KP[a2] / zc;
                for (int m = 0; m < 6; ++m) {
                    double v = 0;
                    for (int k = 0; k < 3; ++k) v += (K[a2 * 3 + k] - (act ? pr * K[6 + k] : 0.0)) / zc * r3[k * 6 + m];
                    Jb[j][a2 * 6 + m] = v;
                }
            }
            for (int m = 0; m < 6; ++m) Jb[j][12 + m] = 0.0;
        }
    }
    double sC[8], sM[8], un[8], u[8][3];
    int goodC, goodM;
    const double prior = rho_cov(Jb, nd, Cm, sC, &goodC);
    const double cov_err = rho_cov(Jb, nd, Mm, sM, &goodM);
    if (!goodC) fl |= 2;
    if (!goodM) fl |= 4;
    double lin = 0;
    for (int j = 0; j < 8; ++j) {
        double n2 = 0;
        for (int r = 0; r < nd; ++r) {
            double v = 0;
            for (int a = 0; a < 6; ++a) v += Jb[j][r * 6 + a] * dth[a];
            u[j][r] = v; n2 += v * v;
        }
        un[j] = sqrt(n2); lin += un[j];
    }
    lin /= 8.0;
    if (loss) *loss = log(prior) + 0.5 * (cov_err + lin) / prior;
    if (flag) *flag = fl;
    if (!gX && !gx && !gs) return;

    /* ---- reverse pass ---- */
    const double g_p = 1.0 / prior - 0.5 * (cov_err + lin) / (prior * prior);
    const double g_c = 0.5 / prior;
    double Cbar[36], Mbar[36], dthbar[6];
    memset(Cbar, 0, sizeof(Cbar)); memset(Mbar, 0, sizeof(Mbar)); memset(dthbar, 0, sizeof(dthbar));
    for (int j = 0; j < 8; ++j) {
        const double wc = goodC ? g_p / (16.0 * sqrt(sC[j])) : 0.0;
        const double wm = goodM ? g_c / (16.0 * sqrt(sM[j])) : 0.0;
        for (int r = 0; r < nd; ++r)
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) {
                    const double qq = Jb[j][r * 6 + a] * Jb[j][r * 6 + b];
                    Cbar[a * 6 + b] += wc * qq;
                    Mbar[a * 6 + b] += wm * qq;
                }
        if (un[j] > 0)
            for (int r = 0; r < nd; ++r)
                for (int a = 0; a < 6; ++a) dthbar[a] += g_c / 8.0 * Jb[j][r * 6 + a] * u[j][r] / un[j];
    }
    double Gbar[36], bbar[6], Hbar[36], T2[36];
    mm6(Cm, Mbar, T1); mm6(T1, Cm, Gbar);                 /* Gbar = C Mbar C */
    for (int a = 0; a < 6; ++a) { double v = 0; for (int k = 0; k < 6; ++k) v += Cm[a * 6 + k] * dthbar[k]; bbar[a] = v; }
    mm6(Mbar, Cm, T1); mm6(T1, G, T2);                    /* Mbar C G */
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) Cbar[a * 6 + b] += T2[a * 6 + b] + T2[b * 6 + a] + dthbar[a] * bvec[b];
    mm6(Cm, Cbar, T1); mm6(T1, Cm, Hbar);
    for (int k = 0; k < 36; ++k) Hbar[k] = -Hbar[k];
    if (fl & 1) memset(Hbar, 0, sizeof(Hbar));            /* torch.where(cond, eye, spd): no gradient into H */

    for (int i = 0; i < N; ++i) {
        double J[12];
        const double* Xi = X + 3 * i;
        point_jac(K, R, t, Xi, J);
        double ecb[2];
        for (int a = 0; a < 2; ++a) {
            const int k = 2 * i + a;
            const double* Jk = J + 6 * a;
            double qh = 0, qg = 0, lb = 0;
            for (int r = 0; r < 6; ++r) {
                double vh = 0, vg = 0;
                for (int c = 0; c < 6; ++c) { vh += Hbar[r * 6 + c] * Jk[c]; vg += Gbar[r * 6 + c] * Jk[c]; }
                qh += Jk[r] * vh; qg += Jk[r] * vg; lb += Jk[r] * bbar[r];
            }
            const double w = W[k], sg = sig[k], sk = s[k];
            const double Wbar = qh + 2.0 * w * sg * qg + ec[k] * lb;
            const double sigbar = w * w * qg;
            const double del = sqrt((q[a] * we) / (sg + 1e-6));
            if (gs) gs[k] = Wbar * (sk > del ? 2.0 * del : 2.0 * sk);
            const double av = fabs(ec[k]);
            const double sgn = (ec[k] > 0) - (ec[k] < 0);
            ecb[a] = sigbar * (av > d[a] ? 2.0 * d[a] : 2.0 * av) * sgn;
            if (gx) gx[k] = ecb[a];
        }
        if (gX) {
            /* dproj/dP = (K[:2,:] - proj (x) K[2,:] [z>0.1]) / max(z,0.1);  gX = -R^T (dproj/dP)^T ecbar */
            double P[3], KPz;
            for (int r = 0; r < 3; ++r) P[r] = R[r * 3] * Xi[0] + R[r * 3 + 1] * Xi[1] + R[r * 3 + 2] * Xi[2] + t[r];
            KPz = K[6] * P[0] + K[7] * P[1] + K[8] * P[2];
            const int act = KPz >= 0.1;
            const double zc = act ? KPz : 0.1;
            double gP[3];
            for (int c = 0; c < 3; ++c) {
                double v = 0;
                for (int a = 0; a < 2; ++a) v += (K[a * 3 + c] - (act ? proj[2 * i + a] * K[6 + c] : 0.0)) / zc * ecb[a];
                gP[c] = v;
            }
            for (int c = 0; c < 3; ++c) gX[3 * i + c] = -(R[0 * 3 + c] * gP[0] + R[1 * 3 + c] * gP[1] + R[2 * 3 + c] * gP[2]);
        }
    }
}

void lc_oracle_pose(const double* K, const double* pose, const double* X, const double* x, const double* s,
                    const double* valid, const double* bbox, int N, double Lmax, double rel, double we,
                    double* loss, double* gX, double* gx, double* gs, double* A, double* C_out, double* M_out,
                    double* Wout, double* sig_out, int* flag, double* scratch) {
    lc_oracle_pose_ex(K, pose, X, x, s, valid, bbox, N, Lmax, rel, we, loss, gX, gx, gs, A, C_out, M_out, Wout, sig_out, flag, scratch, 0);
}

/* Batch driver: contiguous AoS fp64 arrays; OpenMP over poses (mirrors ceres.cpp:161-169 threading). */
void lc_oracle_batch_ex(int B, int N, const double* K, const double* pose, const double* X, const double* x,
                        const double* s, const double* valid, const double* bbox, double Lmax, double rel, double we,
                        double* loss, double* gX, double* gx, double* gs, double* A, double* C, double* M,
                        double* W, double* sig, int* flags, double* scratch /* threads*8*N */, int threads, int cov2d) {
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        const size_t n = (size_t)N;
        lc_oracle_pose_ex(K + 9 * b, pose + 7 * b, X + 3 * n * b, x + 2 * n * b, s + 2 * n * b,
                       valid ? valid + n * b : NULL, bbox + 24 * b, N, Lmax, rel, we,
                       loss ? loss + b : NULL, gX ? gX + 3 * n * b : NULL, gx ? gx + 2 * n * b : NULL,
                       gs ? gs + 2 * n * b : NULL, A ? A + 12 * n * b : NULL, C ? C + 36 * b : NULL,
                       M ? M + 36 * b : NULL, W ? W + 2 * n * b : NULL, sig ? sig + 2 * n * b : NULL,
                       flags ? flags + b : NULL, scratch + (size_t)tid * 8 * n, cov2d);
    }
}

void lc_oracle_batch(int B, int N, const double* K, const double* pose, const double* X, const double* x,
                     const double* s, const double* valid, const double* bbox, double Lmax, double rel, double we,
                     double* loss, double* gX, double* gx, double* gs, double* A, double* C, double* M,
                     double* W, double* sig, int* flags, double* scratch, int threads) {
    lc_oracle_batch_ex(B, N, K, pose, X, x, s, valid, bbox, Lmax, rel, we, loss, gX, gx, gs, A, C, M, W, sig, flags, scratch, threads, 0);
}
