/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU baseline driver for the fused pipeline "P3" of SURVEY.md §8d: per pose,
 *   cer_solver.solve (lib/pnp/cer_solver.py:6-53 -> ceres.cpp:72-145)  [lm_oracle.c]
 *   followed by Loss_cov_mixed(pose := solution) fwd+bwd (lib/cov_mixed.py:100-150)  [lc_oracle.c]
 * one pose per OpenMP thread, fp32 at the boundary like the reference ABI (ext.h:2-15).
 * Used by bench.py's cpu_baseline / --impl reference legs and by tests as the P3 checker.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void lm_oracle_pose(float*, const float*, const float*, const float*, const float*, int, int, float, int, float*, int*,
                    int*, int*, double*, double*);
void lc_oracle_pose(const double*, const double*, const double*, const double*, const double*, const double*,
                    const double*, int, double, double, double, double*, double*, double*, double*, double*, double*,
                    double*, double*, double*, int*, double*);

/* mode bit0: run the LM solve; bit1: run the LC loss fwd+bwd (at the solved pose when bit0, else at `states`) */
void p3_oracle_batch(int B, int N, int mode, float* states /* B x 7 in: start / pose, out: solution */,
                     const float* K, const float* pts3d, const float* pts2d, const float* inv_std, const float* bbox,
                     int max_iter, float ftol, int lm_flags, double Lmax, double rel, double we,
                     float* radius, int* invalid, int* iters, double* loss, float* g_pts3d, float* g_pts2d,
                     float* g_inv_std, int* lc_flags, int threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        const size_t n = (size_t)N;
        float* L = (float*)malloc(sizeof(float) * 4 * n);
        double* d = (double*)malloc(sizeof(double) * (n * (3 + 2 + 2 + 3 + 2 + 2 + 8) + 9 + 7 + 24));
        double *X = d, *x = X + 3 * n, *s = x + 2 * n, *gX = s + 2 * n, *gx = gX + 3 * n, *gs = gx + 2 * n,
               *scr = gs + 2 * n, *Kd = scr + 8 * n, *pd = Kd + 9, *bb = pd + 7;
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            const float* p3 = pts3d + 3 * n * b; const float* p2 = pts2d + 2 * n * b; const float* is = inv_std + 2 * n * b;
            float* st = states + 7 * b;
            if (mode & 1) {
                /* test.py:95 inv_cov = inv_std^2; cer_solver.py:37-38 L = diag(sqrt(icov)) */
                for (size_t i = 0; i < n; ++i) {
                    L[4 * i] = sqrtf(is[2 * i] * is[2 * i]); L[4 * i + 1] = 0; L[4 * i + 2] = 0;
                    L[4 * i + 3] = sqrtf(is[2 * i + 1] * is[2 * i + 1]);
                }
                int it = 0, term = 0;
                lm_oracle_pose(st, K + 9 * b, p2, p3, L, N, max_iter, ftol, lm_flags, radius + b, invalid + b, &it, &term, NULL, NULL);
                if (iters) iters[b] = it;
            }
            if (mode & 2) {
                for (size_t i = 0; i < 3 * n; ++i) X[i] = p3[i];
                for (size_t i = 0; i < 2 * n; ++i) { x[i] = p2[i]; s[i] = is[i]; }
                for (int i = 0; i < 9; ++i) Kd[i] = K[9 * b + i];
                for (int i = 0; i < 7; ++i) pd[i] = st[i];
                for (int i = 0; i < 24; ++i) bb[i] = bbox[24 * b + i];
                int fl = 0;
                lc_oracle_pose(Kd, pd, X, x, s, NULL, bb, N, Lmax, rel, we, loss + b, gX, gx, gs, NULL, NULL, NULL, NULL, NULL, &fl, scr);
                if (lc_flags) lc_flags[b] = fl;
                if (g_pts3d) for (size_t i = 0; i < 3 * n; ++i) g_pts3d[3 * n * b + i] = (float)gX[i];
                if (g_pts2d) for (size_t i = 0; i < 2 * n; ++i) g_pts2d[2 * n * b + i] = (float)gx[i];
                if (g_inv_std) for (size_t i = 0; i < 2 * n; ++i) g_inv_std[2 * n * b + i] = (float)gs[i];
            }
        }
        free(L); free(d);
    }
}
