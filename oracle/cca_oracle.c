/*
 * Plain-C restatement of the criss-cross attention hot path (TEST INFRASTRUCTURE).
 *
 * Follows /root/reference/cc_attention/functions.py:38-47 pixel by pixel, with no
 * tensor library: for pixel (h,w) of sample b the criss-cross set is
 *     column branch: (g, w) for every g != h      (functions.py:38, INF mask at g == h)
 *     row    branch: (h, g) for every g           (functions.py:39)
 * logits = q(:,h,w) . k(:,pos); joint softmax over the H+W-1 entries (functions.py:40);
 * out(:,h,w) = sum_pos A[pos] * v(:,pos)          (functions.py:46-47, out_H + out_W).
 * Backward is the closed form of SURVEY.md 8(a) row a11.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link this.
 * Pinned against the reference's own outputs by tests/test_oracle.py (tests/golden/).
 * Build: gcc -O3 -fopenmp -shared -fPIC -o libcca_oracle.so cca_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX4(b, c, h, w, C, H, W) ((((size_t)(b) * (C) + (c)) * (H) + (h)) * (W) + (w))

/* logits of pixel (b,h,w): e[0..H) column branch (e[h] = -inf), e[H..H+W) row branch */
static void pixel_logits(const double *q, const double *k, int b, int h, int w,
                         int Cq, int H, int W, double *e)
{
    for (int g = 0; g < H; ++g) {
        double s = 0.0;
        for (int c = 0; c < Cq; ++c)
            s += q[IDX4(b, c, h, w, Cq, H, W)] * k[IDX4(b, c, g, w, Cq, H, W)];
        e[g] = (g == h) ? -INFINITY : s;
    }
    for (int g = 0; g < W; ++g) {
        double s = 0.0;
        for (int c = 0; c < Cq; ++c)
            s += q[IDX4(b, c, h, w, Cq, H, W)] * k[IDX4(b, c, h, g, Cq, H, W)];
        e[H + g] = s;
    }
}

static double softmax_inplace(double *e, int n)
{
    double m = -INFINITY, s = 0.0;
    for (int i = 0; i < n; ++i) if (e[i] > m) m = e[i];
    for (int i = 0; i < n; ++i) { e[i] = exp(e[i] - m); s += e[i]; }
    for (int i = 0; i < n; ++i) e[i] /= s;
    return m + log(s);
}

void cca_oracle_forward_f64(const double *q, const double *k, const double *v,
                            double *out, double *lse,
                            int B, int Cq, int C, int H, int W)
{
    const int L = H + W;
#pragma omp parallel
    {
        double *e = (double *)malloc(sizeof(double) * L);
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    pixel_logits(q, k, b, h, w, Cq, H, W, e);
                    lse[((size_t)b * H + h) * W + w] = softmax_inplace(e, L);
                    for (int c = 0; c < C; ++c) {
                        double acc = 0.0;
                        for (int g = 0; g < H; ++g) acc += e[g] * v[IDX4(b, c, g, w, C, H, W)];
                        for (int g = 0; g < W; ++g) acc += e[H + g] * v[IDX4(b, c, h, g, C, H, W)];
                        out[IDX4(b, c, h, w, C, H, W)] = acc;
                    }
                }
        free(e);
    }
}

void cca_oracle_backward_f64(const double *dout, const double *q, const double *k, const double *v,
                             double *dq, double *dk, double *dv,
                             int B, int Cq, int C, int H, int W)
{
    const int L = H + W;
    memset(dq, 0, sizeof(double) * (size_t)B * Cq * H * W);
    memset(dk, 0, sizeof(double) * (size_t)B * Cq * H * W);
    memset(dv, 0, sizeof(double) * (size_t)B * C * H * W);
    /* samples are independent; scatter into dk/dv is sample-local -> parallel over b only */
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        double *a = (double *)malloc(sizeof(double) * L);
        double *da = (double *)malloc(sizeof(double) * L);
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                pixel_logits(q, k, b, h, w, Cq, H, W, a);
                softmax_inplace(a, L);
                double delta = 0.0;
                for (int g = 0; g < H; ++g) {
                    double s = 0.0;
                    for (int c = 0; c < C; ++c)
                        s += dout[IDX4(b, c, h, w, C, H, W)] * v[IDX4(b, c, g, w, C, H, W)];
                    da[g] = s; delta += a[g] * s;
                }
                for (int g = 0; g < W; ++g) {
                    double s = 0.0;
                    for (int c = 0; c < C; ++c)
                        s += dout[IDX4(b, c, h, w, C, H, W)] * v[IDX4(b, c, h, g, C, H, W)];
                    da[H + g] = s; delta += a[H + g] * s;
                }
                for (int g = 0; g < H; ++g) {
                    const double de = a[g] * (da[g] - delta);
                    for (int c = 0; c < C; ++c)
                        dv[IDX4(b, c, g, w, C, H, W)] += a[g] * dout[IDX4(b, c, h, w, C, H, W)];
                    for (int c = 0; c < Cq; ++c) {
                        dq[IDX4(b, c, h, w, Cq, H, W)] += de * k[IDX4(b, c, g, w, Cq, H, W)];
                        dk[IDX4(b, c, g, w, Cq, H, W)] += de * q[IDX4(b, c, h, w, Cq, H, W)];
                    }
                }
                for (int g = 0; g < W; ++g) {
                    const double de = a[H + g] * (da[H + g] - delta);
                    for (int c = 0; c < C; ++c)
                        dv[IDX4(b, c, h, g, C, H, W)] += a[H + g] * dout[IDX4(b, c, h, w, C, H, W)];
                    for (int c = 0; c < Cq; ++c) {
                        dq[IDX4(b, c, h, w, Cq, H, W)] += de * k[IDX4(b, c, h, g, Cq, H, W)];
                        dk[IDX4(b, c, h, g, Cq, H, W)] += de * q[IDX4(b, c, h, w, Cq, H, W)];
                    }
                }
            }
        free(a); free(da);
    }
}
