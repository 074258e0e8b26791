/*
 * oracle/oracle_sparse.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by the
 * product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / the CPU arm).
 *
 * CPU restatement of the reference's Levenberg-Marquardt bundle-adjustment hot path
 * (caomw/BundleAdjustmentMatlab, toolbox/bundle) on an OBSERVATION LIST instead of the
 * reference's dense n x m arrays.  Observations are given in the reference's own
 * traversal order, ascending i + n*j (camera j outer, point i inner;
 * mex_bundle_1_XABeUVWeAeB.c:192-196), so every accumulation below happens in the same
 * order as in the reference and the results are bit-identical to the dense reference
 * wherever the reference's own arithmetic is pinned (invisible cells only ever add an
 * exact zero there: mex_bundle_1_XABeUVWeAeB.c:226-252).  tests/test_oracle.py pins this
 * file against oracle/_ref/libvlgref.so, i.e. the reference's own C compiled unmodified.
 *
 * Parity status: PINNED against outputs of the reference itself run here (oracle/_ref);
 * the reference ships no golden vectors of its own.
 *
 * Each function cites the reference file:line it follows.  Compile with
 * -O2 -ffp-contract=off (see oracle/Makefile); never with -ffast-math / -march=native.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_NA 10

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* VLFeat vl_rodrigues, forward map (third-party, un-vendored; call site
 * reproject_point.h:44).  Same statement as oracle/shim/vl/rodrigues.h. */
void orc_rodrigues(const double *om, double *R)
{
    double th = sqrt(om[0]*om[0] + om[1]*om[1] + om[2]*om[2]);
    if (th < 1e-6) {
        R[0] = 1.0; R[3] = 0.0; R[6] = 0.0;
        R[1] = 0.0; R[4] = 1.0; R[7] = 0.0;
        R[2] = 0.0; R[5] = 0.0; R[8] = 1.0;
        return;
    }
    {
        double x = om[0] / th, y = om[1] / th, z = om[2] / th;
        double xx = x*x, xy = x*y, xz = x*z, yy = y*y, yz = y*z, zz = z*z;
        double sth = sin(th), cth = cos(th), mcth = 1.0 - cth;
        R[0] = 1.0     - mcth*(yy+zz);
        R[1] =   sth*z + mcth*xy;
        R[2] = - sth*y + mcth*xz;
        R[3] = - sth*z + mcth*xy;
        R[4] = 1.0     - mcth*(zz+xx);
        R[5] =   sth*x + mcth*yz;
        R[6] =   sth*y + mcth*xz;
        R[7] = - sth*x + mcth*yz;
        R[8] = 1.0     - mcth*(xx+yy);
    }
}

/* reproject_point.h:16-57.  K9 is the 3x3 column-major matrix the callers build from
 * [fx fy cx cy] (mex_bundle_1_XABeUVWeAeB.c:186-188). */
void orc_reproject(const double *K9, const double *a, const double *b, int nk, double *x)
{
    double K_[9], R[9], Rb[3], x_[3];
    memcpy(K_, K9, sizeof(K_));
    if (nk == 1) {
        K_[0] = a[6]; K_[4] = a[6];
    } else if (nk == 4) {
        K_[0] = a[6]; K_[4] = a[7]; K_[6] = a[8]; K_[7] = a[9];
    }
    orc_rodrigues(a, R);
    Rb[0] = R[0]*b[0] + R[3]*b[1] + R[6]*b[2] + a[3];
    Rb[1] = R[1]*b[0] + R[4]*b[1] + R[7]*b[2] + a[4];
    Rb[2] = R[2]*b[0] + R[5]*b[1] + R[8]*b[2] + a[5];
    x_[0] = K_[0]*Rb[0] + K_[3]*Rb[1] + K_[6]*Rb[2];
    x_[1] = K_[1]*Rb[0] + K_[4]*Rb[1] + K_[7]*Rb[2];
    x_[2] = K_[2]*Rb[0] + K_[5]*Rb[1] + K_[8]*Rb[2];
    x[0] = x_[0] / x_[2];
    x[1] = x_[1] / x_[2];
}

static void make_K9(const double *K4, double *K9)
{
    /* mex_bundle_1_XABeUVWeAeB.c:186-188 */
    K9[0] = K4[0]; K9[3] = 0;     K9[6] = K4[2];
    K9[1] = 0;     K9[4] = K4[1]; K9[7] = K4[3];
    K9[2] = 0;     K9[5] = 0;     K9[8] = 1;
}

/* Per-observation part of mex1 (mex_bundle_1_XABeUVWeAeB.c:196-223) with the two
 * forward-difference helpers (:14-41, :43-70) inlined: h = 1e-10, a1 = a0 + h*da for
 * ALL components (h*0 adds an exact zero), dX = (X1 - X0)/h. */
static void obs_jacobian(const double *K9, const double *a, const double *b, int na,
                         const double *xy, double *X_hat, double *A, double *B, double *e)
{
    const double h = 1e-10;
    int nk = na - 6, k, i;
    double da[ORC_MAX_NA], a1[ORC_MAX_NA], db[3], b1[3], X1[2];
    orc_reproject(K9, a, b, nk, X_hat);
    for (k = 0; k < na; k++) {
        memset(da, 0, sizeof(double) * na);
        da[k] = 1;
        for (i = 0; i < 6 + nk; i++) a1[i] = a[i] + h * da[i];
        orc_reproject(K9, a1, b, nk, X1);
        A[2*k]   = (X1[0] - X_hat[0]) / h;
        A[2*k+1] = (X1[1] - X_hat[1]) / h;
    }
    for (k = 0; k < 3; k++) {
        memset(db, 0, sizeof(db));
        db[k] = 1;
        for (i = 0; i < 3; i++) b1[i] = b[i] + h * db[i];
        orc_reproject(K9, a, b1, nk, X1);
        B[2*k]   = (X1[0] - X_hat[0]) / h;
        B[2*k+1] = (X1[1] - X_hat[1]) / h;
    }
    e[0] = xy[0] - X_hat[0];
    e[1] = xy[1] - X_hat[1];
}

/* CSR by point over a reference-ordered observation list: for point i the positions
 * pt_obs[pt_ptr[i] .. pt_ptr[i+1]) are ascending, i.e. ascending camera j. */
static void build_point_csr(int n, long nobs, const int *obs_pt, long *pt_ptr, long *pt_obs)
{
    long o;
    int i;
    long *fill;
    memset(pt_ptr, 0, sizeof(long) * ((size_t)n + 1));
    for (o = 0; o < nobs; o++) pt_ptr[obs_pt[o] + 1]++;
    for (i = 0; i < n; i++) pt_ptr[i + 1] += pt_ptr[i];
    fill = (long *)malloc(sizeof(long) * ((size_t)n + 1));
    memcpy(fill, pt_ptr, sizeof(long) * ((size_t)n + 1));
    for (o = 0; o < nobs; o++) pt_obs[fill[obs_pt[o]]++] = o;
    free(fill);
}

static void build_cam_ptr(int m, long nobs, const int *obs_cam, long *cam_ptr)
{
    long o;
    int j;
    memset(cam_ptr, 0, sizeof(long) * ((size_t)m + 1));
    for (o = 0; o < nobs; o++) cam_ptr[obs_cam[o] + 1]++;
    for (j = 0; j < m; j++) cam_ptr[j + 1] += cam_ptr[j];
}

/*
 * mex1 on an observation list.  Outputs per observation (list order): X_hat[2], A[2*na]
 * (A[2k+d] = d X_d / d a_k), B[6], e[2], W[na*3] (column-major na x 3); per camera
 * U[na*na] (column-major), eA[na]; per point V[9], eB[3].  Any output pointer except W,
 * U, V, eA, eB may be NULL.  Accumulation order: mex_bundle_1_XABeUVWeAeB.c:266-332.
 */
void orc_stage1(int m, int n, int na, const double *K4, const double *a, const double *b,
                long nobs, const double *obs_xy, const int *obs_pt, const int *obs_cam,
                double *X_hat, double *A, double *B, double *e,
                double *U, double *V, double *W, double *eA, double *eB)
{
    long o;
    int i, j;
    double *Abuf = A ? A : (double *)malloc(sizeof(double) * 2 * na * (size_t)nobs);
    double *Bbuf = B ? B : (double *)malloc(sizeof(double) * 6 * (size_t)nobs);
    double *ebuf = e ? e : (double *)malloc(sizeof(double) * 2 * (size_t)nobs);
    long *cam_ptr = (long *)malloc(sizeof(long) * ((size_t)m + 1));
    long *pt_ptr = (long *)malloc(sizeof(long) * ((size_t)n + 1));
    long *pt_obs = (long *)malloc(sizeof(long) * (size_t)(nobs ? nobs : 1));

    build_cam_ptr(m, nobs, obs_cam, cam_ptr);
    build_point_csr(n, nobs, obs_pt, pt_ptr, pt_obs);

#pragma omp parallel for schedule(static)
    for (o = 0; o < nobs; o++) {
        double K9[9], Xh[2];
        int row, col;
        const double *Ao, *Bo;
        double *Wo = W + (size_t)na * 3 * o;
        make_K9(K4 + 4 * (size_t)obs_cam[o], K9);
        obs_jacobian(K9, a + (size_t)na * obs_cam[o], b + 3 * (size_t)obs_pt[o], na,
                     obs_xy + 2 * o, Xh, Abuf + (size_t)2 * na * o, Bbuf + 6 * o, ebuf + 2 * o);
        if (X_hat) { X_hat[2*o] = Xh[0]; X_hat[2*o+1] = Xh[1]; }
        Ao = Abuf + (size_t)2 * na * o;
        Bo = Bbuf + 6 * o;
        /* W(:,:,i,j) = A' * B into zeroed memory (:305-314) */
        for (col = 0; col < 3; col++)
            for (row = 0; row < na; row++)
                Wo[row + na * col] = 0.0 + (Ao[2*row] * Bo[2*col] + Ao[1+2*row] * Bo[1+2*col]);
    }

    memset(U, 0, sizeof(double) * na * na * (size_t)m);
    memset(eA, 0, sizeof(double) * na * (size_t)m);
    memset(V, 0, sizeof(double) * 9 * (size_t)n);
    memset(eB, 0, sizeof(double) * 3 * (size_t)n);

    /* U_j, eA_j: ascending i for fixed j (:281-290, :317-323) */
#pragma omp parallel for schedule(dynamic, 1)
    for (j = 0; j < m; j++) {
        long p;
        int row, col;
        double *Uj = U + (size_t)na * na * j, *eAj = eA + (size_t)na * j;
        for (p = cam_ptr[j]; p < cam_ptr[j + 1]; p++) {
            const double *Ao = Abuf + (size_t)2 * na * p, *eo = ebuf + 2 * p;
            for (col = 0; col < na; col++)
                for (row = 0; row < na; row++)
                    Uj[row + na * col] += (Ao[2*row] * Ao[2*col] + Ao[1+2*row] * Ao[1+2*col]);
            for (row = 0; row < na; row++)
                eAj[row] += (Ao[2*row] * eo[0] + Ao[1+2*row] * eo[1]);
        }
    }
    /* V_i, eB_i: ascending j for fixed i (:293-302, :326-332) */
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; i++) {
        long q;
        int row, col;
        double *Vi = V + 9 * (size_t)i, *eBi = eB + 3 * (size_t)i;
        for (q = pt_ptr[i]; q < pt_ptr[i + 1]; q++) {
            long p = pt_obs[q];
            const double *Bo = Bbuf + 6 * p, *eo = ebuf + 2 * p;
            for (col = 0; col < 3; col++)
                for (row = 0; row < 3; row++)
                    Vi[row + 3 * col] += (Bo[2*row] * Bo[2*col] + Bo[1+2*row] * Bo[1+2*col]);
            for (row = 0; row < 3; row++)
                eBi[row] += (Bo[2*row] * eo[0] + Bo[1+2*row] * eo[1]);
        }
    }

    if (!A) free(Abuf);
    if (!B) free(Bbuf);
    if (!e) free(ebuf);
    free(cam_ptr); free(pt_ptr); free(pt_obs);
}

/* Y_ij = W_ij * V_inv_i for every observation (bundle_euclid.m:178-184).  MATLAB's
 * matrix product is a BLAS call whose inner order is unspecified; this uses the plain
 * left-to-right triple sum. */
void orc_make_Y(int na, long nobs, const int *obs_pt, const double *W, const double *Vinv, double *Y)
{
    long o;
#pragma omp parallel for schedule(static)
    for (o = 0; o < nobs; o++) {
        const double *Wo = W + (size_t)na * 3 * o, *Vi = Vinv + 9 * (size_t)obs_pt[o];
        double *Yo = Y + (size_t)na * 3 * o;
        int r, c;
        for (c = 0; c < 3; c++)
            for (r = 0; r < na; r++)
                Yo[r + na * c] = Wo[r] * Vi[3*c] + Wo[r + na] * Vi[1 + 3*c] + Wo[r + 2*na] * Vi[2 + 3*c];
    }
}

/*
 * mex2 on an observation list, dense S ((na*m)^2, column-major) and e_ (na*m).
 * S_jk = delta_jk U*_j - sum_i Y_ij W_ik' accumulated over ascending i
 * (mex_bundle_2_Se_.c:72-127); e_j = eA_j - sum_i Y_ij eB_i (:132-155).
 */
void orc_stage2_dense(int m, int n, int na, long nobs, const int *obs_pt, const int *obs_cam,
                      const double *Y, const double *W, const double *U_,
                      const double *eA, const double *eB, double *S, double *e_)
{
    size_t N = (size_t)na * m;
    long *pt_ptr = (long *)malloc(sizeof(long) * ((size_t)n + 1));
    long *pt_obs = (long *)malloc(sizeof(long) * (size_t)(nobs ? nobs : 1));
    long *cam_ptr = (long *)malloc(sizeof(long) * ((size_t)m + 1));
    int i, j, row, col;
    build_point_csr(n, nobs, obs_pt, pt_ptr, pt_obs);
    build_cam_ptr(m, nobs, obs_cam, cam_ptr);
    memset(S, 0, sizeof(double) * N * N);
    for (j = 0; j < m; j++)
        for (col = 0; col < na; col++)
            for (row = 0; row < na; row++)
                S[(row + (size_t)na * j) + N * (col + (size_t)na * j)] = U_[row + na * col + (size_t)na * na * j];
    for (i = 0; i < n; i++) {
        long qj, qk;
        for (qk = pt_ptr[i]; qk < pt_ptr[i + 1]; qk++) {
            const double *Wk = W + (size_t)na * 3 * pt_obs[qk];
            int k = obs_cam[pt_obs[qk]];
            for (qj = pt_ptr[i]; qj < pt_ptr[i + 1]; qj++) {
                const double *Yj = Y + (size_t)na * 3 * pt_obs[qj];
                int jj = obs_cam[pt_obs[qj]];
                for (col = 0; col < na; col++)
                    for (row = 0; row < na; row++)
                        S[(row + (size_t)na * jj) + N * (col + (size_t)na * k)] -= (
                            Yj[row] * Wk[col] + Yj[row + na] * Wk[col + na] + Yj[row + 2*na] * Wk[col + 2*na]);
            }
        }
    }
    for (j = 0; j < m; j++) {
        double YeB[ORC_MAX_NA];
        long p;
        memset(YeB, 0, sizeof(YeB));
        for (p = cam_ptr[j]; p < cam_ptr[j + 1]; p++) {
            const double *Yo = Y + (size_t)na * 3 * p, *eBi = eB + 3 * (size_t)obs_pt[p];
            for (row = 0; row < na; row++)
                YeB[row] += (Yo[row] * eBi[0] + Yo[row + na] * eBi[1] + Yo[row + 2*na] * eBi[2]);
        }
        for (row = 0; row < na; row++) e_[row + (size_t)na * j] = eA[row + (size_t)na * j] - YeB[row];
    }
    free(pt_ptr); free(pt_obs); free(cam_ptr);
}

/*
 * mex3 on an observation list (mex_bundle_3_db_new.c:100-166):
 * db_i = V_inv_i (eB_i - sum_j W_ij' da_j) with ONLY the first six camera parameters
 * entering the sum (:113-120, quirk Q1) unless all_rows != 0; a_new = a + da;
 * b_new = b + db; X_hat_new per observation.
 */
void orc_stage3(int m, int n, int na, const double *K4, const double *a, const double *b,
                long nobs, const int *obs_pt, const int *obs_cam,
                const double *W, const double *da, const double *eB, const double *Vinv,
                int all_rows, double *db, double *a_new, double *b_new, double *X_hat_new)
{
    long *pt_ptr = (long *)malloc(sizeof(long) * ((size_t)n + 1));
    long *pt_obs = (long *)malloc(sizeof(long) * (size_t)(nobs ? nobs : 1));
    int i, nrow = all_rows ? na : 6;
    long o;
    size_t t;
    build_point_csr(n, nobs, obs_pt, pt_ptr, pt_obs);
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; i++) {
        double Wda[3];
        long q;
        int row, r;
        Wda[0] = eB[3*(size_t)i]; Wda[1] = eB[1+3*(size_t)i]; Wda[2] = eB[2+3*(size_t)i];
        for (q = pt_ptr[i]; q < pt_ptr[i + 1]; q++) {
            long p = pt_obs[q];
            const double *daj = da + (size_t)na * obs_cam[p];
            for (row = 0; row < 3; row++) {
                const double *Wc = W + (size_t)na * 3 * p + (size_t)na * row;
                double s = Wc[0] * daj[0];
                for (r = 1; r < nrow; r++) s = s + Wc[r] * daj[r];
                Wda[row] -= s;
            }
        }
        for (row = 0; row < 3; row++)
            db[row + 3*(size_t)i] = (Vinv[row + 9*(size_t)i] * Wda[0] + Vinv[row + 3 + 9*(size_t)i] * Wda[1]
                                     + Vinv[row + 6 + 9*(size_t)i] * Wda[2]);
    }
    for (t = 0; t < (size_t)na * m; t++) a_new[t] = a[t] + da[t];
    for (t = 0; t < 3 * (size_t)n; t++) b_new[t] = b[t] + db[t];
    if (X_hat_new) {
#pragma omp parallel for schedule(static)
        for (o = 0; o < nobs; o++) {
            double K9[9];
            make_K9(K4 + 4 * (size_t)obs_cam[o], K9);
            orc_reproject(K9, a_new + (size_t)na * obs_cam[o], b_new + 3 * (size_t)obs_pt[o], na - 6,
                          X_hat_new + 2 * o);
        }
    }
    free(pt_ptr); free(pt_obs);
}

/* sum of squared residuals X - X_hat over the list (bundle_euclid.m:205-210; the
 * reference's e'*e is a BLAS ddot of unspecified order -> tolerance-based everywhere). */
double orc_cost(long nobs, const double *obs_xy, const double *X_hat)
{
    double s = 0.0;
    long o;
#pragma omp parallel for reduction(+:s) schedule(static)
    for (o = 0; o < 2 * nobs; o++) {
        double d = obs_xy[o] - X_hat[o];
        s += d * d;
    }
    return s;
}

/* ------------------------------------------------------------------------------------
 * CPU arm for the large configurations (bench.py cpu_baseline / --impl reference).
 *
 * The reference itself cannot run them: its arrays are dense n x m (W alone is
 * 144*n*m bytes, mex_bundle_1_XABeUVWeAeB.c:165) and its solve is pinv of a dense S
 * (bundle_euclid.m:193).  This is therefore a PORT of the same LM trial step
 * (bundle_euclid.m:139-241) with the reduced system solved by block-Jacobi
 * preconditioned CG on the implicit Schur complement, threaded with OpenMP.
 * ------------------------------------------------------------------------------------ */

/* pinv of a symmetric PSD k x k block via Cholesky with elimination of exactly-zero
 * pivots (pinv(0-row/col) = 0, bundle_euclid.m:180,193 semantics for structural zeros). */
static void sym_pinv_small(int k, const double *M, double *Minv)
{
    double L[ORC_MAX_NA * ORC_MAX_NA], Li[ORC_MAX_NA * ORC_MAX_NA];
    int i, j, r;
    memset(L, 0, sizeof(L));
    memset(Li, 0, sizeof(Li));
    for (j = 0; j < k; j++) {
        double d = M[j + k * j];
        for (r = 0; r < j; r++) d -= L[j + k * r] * L[j + k * r];
        if (d > 0.0) {
            double inv;
            L[j + k * j] = sqrt(d);
            inv = 1.0 / L[j + k * j];
            for (i = j + 1; i < k; i++) {
                double s = M[i + k * j];
                for (r = 0; r < j; r++) s -= L[i + k * r] * L[j + k * r];
                L[i + k * j] = s * inv;
            }
        } else {
            L[j + k * j] = 0.0;
            for (i = j + 1; i < k; i++) L[i + k * j] = 0.0;
        }
    }
    /* Li = L^-1 restricted to non-eliminated rows */
    for (j = 0; j < k; j++) {
        if (L[j + k * j] == 0.0) continue;
        Li[j + k * j] = 1.0 / L[j + k * j];
        for (i = j + 1; i < k; i++) {
            double s = 0.0;
            if (L[i + k * i] == 0.0) continue;
            for (r = j; r < i; r++) s -= L[i + k * r] * Li[r + k * j];
            Li[i + k * j] = s / L[i + k * i];
        }
    }
    for (j = 0; j < k; j++)
        for (i = 0; i < k; i++) {
            double s = 0.0;
            for (r = (i > j ? i : j); r < k; r++) s += Li[r + k * i] * Li[r + k * j];
            Minv[i + k * j] = s;
        }
}

typedef struct {
    int m, n, na;
    long nobs;
    const int *obs_pt, *obs_cam;
    const long *cam_ptr, *pt_ptr, *pt_obs;
    const double *W, *Vinv, *Ud;
} schur_op;

/* q = S p = U* p - W V*^-1 W' p, two sweeps over the observation list */
static void schur_matvec(const schur_op *s, const double *p, double *q, double *t)
{
    int i, j, na = s->na;
#pragma omp parallel for schedule(static)
    for (i = 0; i < s->n; i++) {
        double acc[3] = {0, 0, 0};
        long k;
        int r, c;
        const double *Vi = s->Vinv + 9 * (size_t)i;
        for (k = s->pt_ptr[i]; k < s->pt_ptr[i + 1]; k++) {
            long o = s->pt_obs[k];
            const double *Wo = s->W + (size_t)na * 3 * o, *pj = p + (size_t)na * s->obs_cam[o];
            for (c = 0; c < 3; c++)
                for (r = 0; r < na; r++) acc[c] += Wo[r + na * c] * pj[r];
        }
        for (r = 0; r < 3; r++) t[r + 3*(size_t)i] = Vi[r] * acc[0] + Vi[r + 3] * acc[1] + Vi[r + 6] * acc[2];
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (j = 0; j < s->m; j++) {
        double acc[ORC_MAX_NA];
        long o;
        int r, c;
        const double *Uj = s->Ud + (size_t)na * na * j, *pj = p + (size_t)na * j;
        for (r = 0; r < na; r++) {
            double v = 0.0;
            for (c = 0; c < na; c++) v += Uj[r + na * c] * pj[c];
            acc[r] = v;
        }
        for (o = s->cam_ptr[j]; o < s->cam_ptr[j + 1]; o++) {
            const double *Wo = s->W + (size_t)na * 3 * o, *ti = t + 3 * (size_t)s->obs_pt[o];
            for (r = 0; r < na; r++) acc[r] -= Wo[r] * ti[0] + Wo[r + na] * ti[1] + Wo[r + 2*na] * ti[2];
        }
        for (r = 0; r < na; r++) q[r + (size_t)na * j] = acc[r];
    }
}

/*
 * One LM trial step (bundle_euclid.m:139-217) on an observation list with a PCG solve.
 * In: K4, a, b, lambda.  Out: a_new, b_new, costs[0]=old, costs[1]=new,
 * costs[2]=dp'(lambda dp + g), stage_seconds[0..3] = stage1 / vinv+precond / pcg / stage3;
 * returns the PCG iteration count.
 */
int orc_trial_step_pcg(int m, int n, int na, const double *K4, const double *a, const double *b,
                       long nobs, const double *obs_xy, const int *obs_pt, const int *obs_cam,
                       double lambda, double pcg_rtol, int pcg_max_iter,
                       double *a_new, double *b_new, double *costs, double *stage_seconds)
{
    size_t N = (size_t)na * m;
    double *X_hat = (double *)malloc(sizeof(double) * 2 * (size_t)nobs);
    double *U = (double *)malloc(sizeof(double) * na * N), *eA = (double *)malloc(sizeof(double) * N);
    double *V = (double *)malloc(sizeof(double) * 9 * (size_t)n), *eB = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double *W = (double *)malloc(sizeof(double) * 3 * na * (size_t)nobs);
    double *Vinv = (double *)malloc(sizeof(double) * 9 * (size_t)n);
    double *Minv = (double *)malloc(sizeof(double) * na * N);
    double *e_ = (double *)malloc(sizeof(double) * N), *da = (double *)calloc(N, sizeof(double));
    double *r = (double *)malloc(sizeof(double) * N), *z = (double *)malloc(sizeof(double) * N);
    double *p = (double *)malloc(sizeof(double) * N), *q = (double *)malloc(sizeof(double) * N);
    double *t = (double *)malloc(sizeof(double) * 3 * (size_t)n), *db = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    long *pt_ptr = (long *)malloc(sizeof(long) * ((size_t)n + 1)), *pt_obs = (long *)malloc(sizeof(long) * (size_t)nobs);
    long *cam_ptr = (long *)malloc(sizeof(long) * ((size_t)m + 1));
    schur_op op;
    int i, j, it = 0;
    double t0, t1, rz, r0norm, denom = 0.0;
    size_t k;
#ifdef _OPENMP
#define ORC_NOW() omp_get_wtime()
#else
#define ORC_NOW() 0.0
#endif
    build_point_csr(n, nobs, obs_pt, pt_ptr, pt_obs);
    build_cam_ptr(m, nobs, obs_cam, cam_ptr);

    t0 = ORC_NOW();
    orc_stage1(m, n, na, K4, a, b, nobs, obs_xy, obs_pt, obs_cam, X_hat, NULL, NULL, NULL, U, V, W, eA, eB);
    costs[0] = orc_cost(nobs, obs_xy, X_hat);
    t1 = ORC_NOW(); stage_seconds[0] = t1 - t0; t0 = t1;

    /* damping (bundle_euclid.m:162-173) and V*^-1 (:178-181) */
    for (j = 0; j < m; j++)
        for (i = 0; i < na; i++) U[i + na * i + (size_t)na * na * j] = (1 + lambda) * U[i + na * i + (size_t)na * na * j];
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; i++) {
        double Vd[9];
        int d;
        memcpy(Vd, V + 9 * (size_t)i, sizeof(Vd));
        for (d = 0; d < 3; d++) Vd[4 * d] = (1 + lambda) * Vd[4 * d];
        sym_pinv_small(3, Vd, Vinv + 9 * (size_t)i);
    }
    /* block-Jacobi blocks S_jj and e_ */
#pragma omp parallel for schedule(dynamic, 4)
    for (j = 0; j < m; j++) {
        double Sjj[ORC_MAX_NA * ORC_MAX_NA], ej[ORC_MAX_NA], Yo[3 * ORC_MAX_NA];
        long o;
        int rr, c, d;
        memcpy(Sjj, U + (size_t)na * na * j, sizeof(double) * na * na);
        memcpy(ej, eA + (size_t)na * j, sizeof(double) * na);
        for (o = cam_ptr[j]; o < cam_ptr[j + 1]; o++) {
            const double *Wo = W + (size_t)na * 3 * o, *Vi = Vinv + 9 * (size_t)obs_pt[o], *eBi = eB + 3 * (size_t)obs_pt[o];
            for (c = 0; c < 3; c++)
                for (rr = 0; rr < na; rr++)
                    Yo[rr + na * c] = Wo[rr] * Vi[3*c] + Wo[rr + na] * Vi[1 + 3*c] + Wo[rr + 2*na] * Vi[2 + 3*c];
            for (c = 0; c < na; c++)
                for (rr = 0; rr < na; rr++)
                    for (d = 0; d < 3; d++) Sjj[rr + na * c] -= Yo[rr + na * d] * Wo[c + na * d];
            for (rr = 0; rr < na; rr++) ej[rr] -= Yo[rr] * eBi[0] + Yo[rr + na] * eBi[1] + Yo[rr + 2*na] * eBi[2];
        }
        sym_pinv_small(na, Sjj, Minv + (size_t)na * na * j);
        memcpy(e_ + (size_t)na * j, ej, sizeof(double) * na);
    }
    t1 = ORC_NOW(); stage_seconds[1] = t1 - t0; t0 = t1;

    op.m = m; op.n = n; op.na = na; op.nobs = nobs; op.obs_pt = obs_pt; op.obs_cam = obs_cam;
    op.cam_ptr = cam_ptr; op.pt_ptr = pt_ptr; op.pt_obs = pt_obs; op.W = W; op.Vinv = Vinv; op.Ud = U;
    memcpy(r, e_, sizeof(double) * N);
    r0norm = 0.0;
    for (k = 0; k < N; k++) r0norm += r[k] * r[k];
    r0norm = sqrt(r0norm);
    rz = 0.0;
    for (j = 0; j < m; j++)
        for (i = 0; i < na; i++) {
            double v = 0.0;
            int c;
            for (c = 0; c < na; c++) v += Minv[i + na * c + (size_t)na * na * j] * r[c + (size_t)na * j];
            z[i + (size_t)na * j] = v;
            rz += v * r[i + (size_t)na * j];
        }
    memcpy(p, z, sizeof(double) * N);
    if (r0norm > 0.0) {
        for (it = 0; it < pcg_max_iter; ) {
            double pq = 0.0, alpha, rz_new = 0.0, rn = 0.0, beta;
            schur_matvec(&op, p, q, t);
            for (k = 0; k < N; k++) pq += p[k] * q[k];
            if (!(pq > 0.0)) break;
            alpha = rz / pq;
            for (k = 0; k < N; k++) { da[k] += alpha * p[k]; r[k] -= alpha * q[k]; rn += r[k] * r[k]; }
            it++;
            if (sqrt(rn) <= pcg_rtol * r0norm) break;
            for (j = 0; j < m; j++)
                for (i = 0; i < na; i++) {
                    double v = 0.0;
                    int c;
                    for (c = 0; c < na; c++) v += Minv[i + na * c + (size_t)na * na * j] * r[c + (size_t)na * j];
                    z[i + (size_t)na * j] = v;
                    rz_new += v * r[i + (size_t)na * j];
                }
            beta = rz_new / rz;
            rz = rz_new;
            for (k = 0; k < N; k++) p[k] = z[k] + beta * p[k];
        }
    }
    t1 = ORC_NOW(); stage_seconds[2] = t1 - t0; t0 = t1;

    orc_stage3(m, n, na, K4, a, b, nobs, obs_pt, obs_cam, W, da, eB, Vinv, 0, db, a_new, b_new, X_hat);
    costs[1] = orc_cost(nobs, obs_xy, X_hat);
    for (k = 0; k < N; k++) denom += da[k] * (lambda * da[k] + eA[k]);
    for (k = 0; k < 3 * (size_t)n; k++) denom += db[k] * (lambda * db[k] + eB[k]);
    costs[2] = denom;
    t1 = ORC_NOW(); stage_seconds[3] = t1 - t0;

    free(X_hat); free(U); free(eA); free(V); free(eB); free(W); free(Vinv); free(Minv); free(e_); free(da);
    free(r); free(z); free(p); free(q); free(t); free(db); free(pt_ptr); free(pt_obs); free(cam_ptr);
    return it;
}
