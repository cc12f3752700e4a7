/*
 * oracle/srps_oracle_port.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Plain-C (OpenMP) fp32 transcription of one SRmeetsPS outer iteration on the reference's
 * masked-vector layouts.  It restates nihalsid/SRmeetsPS-CUDA:
 *     lighting   SRmeetsPS-GPU/devicecalls.cu:376-444   (4x4 normal equations + warm CG)
 *     albedo     devicecalls.cu:497-548                 (CG on the diagonal system)
 *     depth      devicecalls.cu:550-786                 (KtK + AtA solve, 101-pass CG, energy)
 *     CG         devicecalls.cu:229-279
 *     normals    devicecalls.cu:171-223, SRPS.cu:310-315
 * with the sparse products replaced by the equivalent per-pixel stencil form,
 * which tests/test_oracle.py proves equal to the assembled matrices of
 * oracle/srps_oracle.py.  Nothing in the product path (srmeetsps-cuda_b200/) links, loads
 * or calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.  Parity pin: see the header of oracle/srps_oracle.py.
 *
 * Geometry comes in as index arrays built by oracle/port.py from the reference-style
 * operators (SRPS.cu:23-71 make_gradient; Utilities.cpp:201-220 + SRPS.cu:172-193 KT):
 *   (Dx v)_p = dx_sg[p] * (v[dx_nb[p]] - v[p])   sg=+1 forward, -1 backward, 0 none
 *   rows of Dx that touch column p besides row p: x_prev_f[p] (a forward row, +1) and
 *   x_next_b[p] (a backward row, -1); -1 when absent.  Same for y.
 *   kt_idx[b][0..sf*sf) = HR masked indices averaged by LR masked pixel b; lr_of[p] = b or -1.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int npix, npixs, n, c, sf;
    float fx, fy;
    const int *dx_nb, *dy_nb;
    const float *dx_sg, *dy_sg;
    const int *x_prev_f, *x_next_b, *y_prev_f, *y_next_b;
    const int *kt_idx;
    const int *lr_of;
    const float *xx, *yy;
} srps_geom;

int srps_port_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- normals: devicecalls.cu:171-223 --------------------------------------------------- */
void srps_port_normals(const srps_geom *g, const float *z, float *N, float *dz) {
    const int P = g->npix;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; p++) {
        float zx = g->dx_sg[p] * (z[g->dx_nb[p]] - z[p]);
        float zy = g->dy_sg[p] * (z[g->dy_nb[p]] - z[p]);
        float n0 = g->fx * zx, n1 = g->fy * zy;
        float n2 = -z[p] - g->xx[p] * zx - g->yy[p] * zy;
        float d = fmaxf(1e-10f, sqrtf(n0 * n0 + n1 * n1 + n2 * n2));
        N[p] = n0 / d; N[P + p] = n1 / d; N[2 * P + p] = n2 / d; N[3 * P + p] = 1.f;
        dz[p] = d;
    }
}

/* ---- reference CG on a tiny dense SPD matrix (lighting): devicecalls.cu:229-279 -------- */
static int cg_dense4(const float A[16], float x[4], float b[4]) {
    const float tol2 = 1e-9f * 1e-9f;
    float p[4] = {0, 0, 0, 0}, om[4];
    float r0 = 0.f, r1 = 0.f;
    int k = 0;
    for (int i = 0; i < 4; i++) r1 += b[i] * b[i];
    while (r1 > tol2 && k <= 100) {
        k++;
        if (k == 1) { for (int i = 0; i < 4; i++) p[i] = b[i]; }
        else { float beta = r1 / r0; for (int i = 0; i < 4; i++) p[i] = beta * p[i] + b[i]; }
        float dot = 0.f;
        for (int i = 0; i < 4; i++) {
            om[i] = 0.f;
            for (int j = 0; j < 4; j++) om[i] += A[i * 4 + j] * p[j];
            dot += p[i] * om[i];
        }
        float alpha = r1 / dot;
        for (int i = 0; i < 4; i++) { x[i] += alpha * p[i]; b[i] -= alpha * om[i]; }
        r0 = r1; r1 = 0.f;
        for (int i = 0; i < 4; i++) r1 += b[i] * b[i];
    }
    return k;
}

/* ---- lighting: devicecalls.cu:408-444 -------------------------------------------------- */
void srps_port_lighting(const srps_geom *g, float *s, const float *rho, const float *N, const float *I) {
    const int P = g->npix, n = g->n, C = g->c;
    for (int ch = 0; ch < C; ch++) {
        double ata[16];
        memset(ata, 0, sizeof ata);
        double *atb = (double *)calloc((size_t)n * 4, sizeof(double));
#pragma omp parallel
        {
            double la[10];
            memset(la, 0, sizeof la);
            double *lb = (double *)calloc((size_t)n * 4, sizeof(double));
#pragma omp for schedule(static)
            for (int p0 = 0; p0 < P; p0 += 512) {
                int p1 = p0 + 512 < P ? p0 + 512 : P;
                for (int p = p0; p < p1; p++) {
                    float r = rho[(size_t)ch * P + p];
                    float a[4] = {r * N[p], r * N[P + p], r * N[2 * (size_t)P + p], r * N[3 * (size_t)P + p]};
                    int q = 0;
                    for (int i = 0; i < 4; i++) for (int j = i; j < 4; j++) la[q++] += (double)(a[i] * a[j]);
                }
                for (int im = 0; im < n; im++) {
                    const float *Ip = I + ((size_t)im * C + ch) * P;
                    float b0 = 0, b1 = 0, b2 = 0, b3 = 0;
                    for (int p = p0; p < p1; p++) {
                        float r = rho[(size_t)ch * P + p], v = Ip[p];
                        b0 += r * N[p] * v; b1 += r * N[P + p] * v;
                        b2 += r * N[2 * (size_t)P + p] * v; b3 += r * N[3 * (size_t)P + p] * v;
                    }
                    lb[im * 4 + 0] += b0; lb[im * 4 + 1] += b1; lb[im * 4 + 2] += b2; lb[im * 4 + 3] += b3;
                }
            }
#pragma omp critical
            {
                int q = 0;
                for (int i = 0; i < 4; i++) for (int j = i; j < 4; j++) { ata[i * 4 + j] += la[q]; if (i != j) ata[j * 4 + i] += la[q]; q++; }
                for (int t = 0; t < n * 4; t++) atb[t] += lb[t];
            }
            free(lb);
        }
        float A[16];
        for (int t = 0; t < 16; t++) A[t] = (float)ata[t];
        for (int im = 0; im < n; im++) {
            float *x = s + ((size_t)im * C + ch) * 4;
            float b[4];
            for (int i = 0; i < 4; i++) {                      /* residual: devicecalls.cu:424 */
                float ax = 0.f;
                for (int j = 0; j < 4; j++) ax += A[i * 4 + j] * x[j];
                b[i] = (float)atb[im * 4 + i] - ax;
            }
            cg_dense4(A, x, b);                                /* devicecalls.cu:437 */
        }
        free(atb);
    }
}

/* ---- generic reference CG over masked vectors: devicecalls.cu:229-279 ------------------ */
typedef void (*matvec_fn)(const void *ctx, const float *p, float *out);

static int cg_vec(int N, matvec_fn mv, const void *ctx, float *x, float *b, float *p, float *om) {
    const double tol2 = (double)(1e-9f * 1e-9f);
    double r0 = 0, r1 = 0;
    int k = 0;
#pragma omp parallel for reduction(+ : r1) schedule(static)
    for (int i = 0; i < N; i++) r1 += (double)(b[i] * b[i]);
    while ((float)r1 > (float)tol2 && k <= 100) {
        k++;
        if (k == 1) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < N; i++) p[i] = b[i];
        } else {
            float beta = (float)r1 / (float)r0;
#pragma omp parallel for schedule(static)
            for (int i = 0; i < N; i++) p[i] = beta * p[i] + b[i];
        }
        mv(ctx, p, om);
        double dot = 0;
#pragma omp parallel for reduction(+ : dot) schedule(static)
        for (int i = 0; i < N; i++) dot += (double)(p[i] * om[i]);
        float alpha = (float)r1 / (float)dot;
        r0 = r1; r1 = 0;
#pragma omp parallel for reduction(+ : r1) schedule(static)
        for (int i = 0; i < N; i++) {
            x[i] += alpha * p[i];
            b[i] -= alpha * om[i];
            r1 += (double)(b[i] * b[i]);
        }
    }
    return k;
}

/* ---- albedo: devicecalls.cu:497-548 ---------------------------------------------------- */
typedef struct { const float *d; int P; } diag_ctx;
static void diag_mv(const void *c, const float *p, float *out) {
    const diag_ctx *d = (const diag_ctx *)c;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < d->P; i++) out[i] = d->d[i] * p[i];
}

void srps_port_albedo(const srps_geom *g, const float *s, float *rho, const float *N, const float *I,
                      int closed_form, int *iters_out) {
    const int P = g->npix, n = g->n, C = g->c;
    float *d = (float *)malloc(sizeof(float) * P), *b = (float *)malloc(sizeof(float) * P);
    float *pv = (float *)malloc(sizeof(float) * P), *om = (float *)malloc(sizeof(float) * P);
    for (int ch = 0; ch < C; ch++) {
        float *r = rho + (size_t)ch * P;
#pragma omp parallel for schedule(static)
        for (int p = 0; p < P; p++) {
            float dd = 0.f, bb = 0.f;
            float n0 = N[p], n1 = N[P + p], n2 = N[2 * (size_t)P + p], n3 = N[3 * (size_t)P + p];
            for (int im = 0; im < n; im++) {
                const float *sv = s + ((size_t)im * C + ch) * 4;
                float a = n0 * sv[0] + n1 * sv[1] + n2 * sv[2] + n3 * sv[3];   /* devicecalls.cu:507 */
                dd += a * a;
                bb += a * I[((size_t)im * C + ch) * P + p];
            }
            d[p] = dd;
            b[p] = closed_form ? bb : bb - dd * r[p];                             /* :404-405 */
        }
        if (closed_form) {
#pragma omp parallel for schedule(static)
            for (int p = 0; p < P; p++) if (d[p] > 0.f) r[p] = b[p] / d[p];
            if (iters_out) iters_out[ch] = 0;
        } else {
            diag_ctx dc = {d, P};
            int k = cg_vec(P, diag_mv, &dc, r, b, pv, om);                        /* :540 */
            if (iters_out) iters_out[ch] = k;
        }
    }
    free(d); free(b); free(pv); free(om);
}

/* ---- depth: devicecalls.cu:550-786 ----------------------------------------------------- */
typedef struct {
    const srps_geom *g;
    const float *M;     /* [6][P]: m00 m01 m02 m11 m12 m22 */
    float *q0, *q1, *ks;
} depth_ctx;

static void depth_mv(const void *c, const float *v, float *out) {
    const depth_ctx *d = (const depth_ctx *)c;
    const srps_geom *g = d->g;
    const int P = g->npix, S = g->npixs, bs = g->sf * g->sf;
    const float *M = d->M;
    const float inv = 1.f / (float)bs;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < S; b++) {
        float acc = 0.f;
        for (int t = 0; t < bs; t++) acc += v[g->kt_idx[(size_t)b * bs + t]];
        d->ks[b] = acc * inv;                                     /* K v */
    }
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; p++) {
        float gx = g->dx_sg[p] * (v[g->dx_nb[p]] - v[p]);
        float gy = g->dy_sg[p] * (v[g->dy_nb[p]] - v[p]);
        float gz = v[p];
        d->q0[p] = M[p] * gx + M[P + p] * gy + M[2 * (size_t)P + p] * gz;
        d->q1[p] = M[P + p] * gx + M[3 * (size_t)P + p] * gy + M[4 * (size_t)P + p] * gz;
    }
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; p++) {
        float gx = g->dx_sg[p] * (v[g->dx_nb[p]] - v[p]);
        float gy = g->dy_sg[p] * (v[g->dy_nb[p]] - v[p]);
        float y = M[2 * (size_t)P + p] * gx + M[4 * (size_t)P + p] * gy + M[5 * (size_t)P + p] * v[p];
        y += -g->dx_sg[p] * d->q0[p] - g->dy_sg[p] * d->q1[p];
        if (g->x_prev_f[p] >= 0) y += d->q0[g->x_prev_f[p]];
        if (g->x_next_b[p] >= 0) y -= d->q0[g->x_next_b[p]];
        if (g->y_prev_f[p] >= 0) y += d->q1[g->y_prev_f[p]];
        if (g->y_next_b[p] >= 0) y -= d->q1[g->y_next_b[p]];
        if (g->lr_of[p] >= 0) y += d->ks[g->lr_of[p]] * inv;      /* Kt K v */
        out[p] = y;
    }
}

/* returns the energy ||Kz-z0s||^2 + ||Az-B||^2 with lagged A,B and the new z (:762-767,785) */
float srps_port_depth(const srps_geom *g, const float *s, const float *rho, const float *I,
                      const float *dz, const float *z0s, float *z, int *iters_out) {
    const int P = g->npix, S = g->npixs, n = g->n, C = g->c, bs = g->sf * g->sf;
    float *M = (float *)malloc(sizeof(float) * 6 * (size_t)P);
    float *gv = (float *)malloc(sizeof(float) * 3 * (size_t)P);
    float *e0 = (float *)malloc(sizeof(float) * P);
    float *q0 = (float *)malloc(sizeof(float) * P), *q1 = (float *)malloc(sizeof(float) * P);
    float *ks = (float *)malloc(sizeof(float) * (S > 0 ? S : 1));
    float *r = (float *)malloc(sizeof(float) * P), *pv = (float *)malloc(sizeof(float) * P);
    float *om = (float *)malloc(sizeof(float) * P);
    /* rows t_{c,j}(p) and B (devicecalls.cu:550-620) folded into M = sum t t^T, g = sum t B, e0 = sum B^2 */
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; p++) {
        float m[6] = {0, 0, 0, 0, 0, 0}, gg[3] = {0, 0, 0}, ee = 0.f;
        for (int ch = 0; ch < C; ch++) {
            float rr = rho[(size_t)ch * P + p], rd = rr / dz[p];
            for (int im = 0; im < n; im++) {
                const float *sv = s + ((size_t)im * C + ch) * 4;
                float t0 = rd * (g->fx * sv[0] - g->xx[p] * sv[2]);
                float t1 = rd * (g->fy * sv[1] - g->yy[p] * sv[2]);
                float t2 = -rd * sv[2];
                float B = I[((size_t)im * C + ch) * P + p] - rr * sv[3];
                m[0] += t0 * t0; m[1] += t0 * t1; m[2] += t0 * t2;
                m[3] += t1 * t1; m[4] += t1 * t2; m[5] += t2 * t2;
                gg[0] += t0 * B; gg[1] += t1 * B; gg[2] += t2 * B;
                ee += B * B;
            }
        }
        for (int t = 0; t < 6; t++) M[(size_t)t * P + p] = m[t];
        for (int t = 0; t < 3; t++) gv[(size_t)t * P + p] = gg[t];
        e0[p] = ee;
    }
    depth_ctx dc = {g, M, q0, q1, ks};
    /* residual r = Kt z0s + G^T g - A_ z  (devicecalls.cu:743-745,758) */
    depth_mv(&dc, z, om);
    const float inv = 1.f / (float)bs;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; p++) {
        float rhs = gv[2 * (size_t)P + p] - g->dx_sg[p] * gv[p] - g->dy_sg[p] * gv[P + p];
        if (g->x_prev_f[p] >= 0) rhs += gv[g->x_prev_f[p]];
        if (g->x_next_b[p] >= 0) rhs -= gv[g->x_next_b[p]];
        if (g->y_prev_f[p] >= 0) rhs += gv[P + g->y_prev_f[p]];
        if (g->y_next_b[p] >= 0) rhs -= gv[P + g->y_next_b[p]];
        if (g->lr_of[p] >= 0) rhs += z0s[g->lr_of[p]] * inv;
        r[p] = rhs - om[p];
    }
    int k = cg_vec(P, depth_mv, &dc, z, r, pv, om);               /* devicecalls.cu:759 */
    if (iters_out) *iters_out = k;
    double e_d = 0, e_p = 0;
#pragma omp parallel for reduction(+ : e_d) schedule(static)
    for (int b = 0; b < S; b++) {
        float acc = 0.f;
        for (int t = 0; t < bs; t++) acc += z[g->kt_idx[(size_t)b * bs + t]];
        float df = acc * inv - z0s[b];
        e_d += (double)(df * df);
    }
#pragma omp parallel for reduction(+ : e_p) schedule(static)
    for (int p = 0; p < P; p++) {
        double gx = g->dx_sg[p] * (z[g->dx_nb[p]] - z[p]);
        double gy = g->dy_sg[p] * (z[g->dy_nb[p]] - z[p]);
        double gz = z[p];
        double m0 = M[p], m1 = M[P + p], m2 = M[2 * (size_t)P + p], m3 = M[3 * (size_t)P + p],
               m4 = M[4 * (size_t)P + p], m5 = M[5 * (size_t)P + p];
        double quad = gx * (m0 * gx + m1 * gy + m2 * gz) + gy * (m1 * gx + m3 * gy + m4 * gz) +
                      gz * (m2 * gx + m4 * gy + m5 * gz);
        double lin = gv[p] * gx + gv[P + p] * gy + gv[2 * (size_t)P + p] * gz;
        e_p += quad - 2.0 * lin + (double)e0[p];
    }
    free(M); free(gv); free(e0); free(q0); free(q1); free(ks); free(r); free(pv); free(om);
    return (float)(e_d + e_p);
}

/* ---- one pass of the do-while body: SRPS.cu:276-317 ------------------------------------ */
float srps_port_outer_iteration(const srps_geom *g, float *s, float *rho, float *z, float *N, float *dz,
                                const float *I, const float *z0s, int albedo_closed_form,
                                int *depth_iters, int *albedo_iters) {
    srps_port_lighting(g, s, rho, N, I);
    srps_port_albedo(g, s, rho, N, I, albedo_closed_form, albedo_iters);
    float e = srps_port_depth(g, s, rho, I, dz, z0s, z, depth_iters);
    srps_port_normals(g, z, N, dz);
    return e;
}
