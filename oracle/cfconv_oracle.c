/*
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * CPU restatement of the SchNet continuous-filter convolution of the reference (paths relative to /root/reference/src/schnet/):
 *   - half neighbour list, strict r2 < rc2 ............ CpuCFConv.cpp:94-116 (minimum image :30-55)
 *   - forward (Gaussians -> dense -> act -> dense -> cutoff, symmetric scatter) ... CpuCFConv.cpp:133-188
 *   - backward (input and position gradients) ........ CpuCFConv.cpp:211-299
 *   - cosine cutoff ................................... CpuCFConv.cpp:301-307
 * w1 is indexed [width][numGaussians] row-major and w2 [width][width], exactly as the reference indexes the memory it is handed
 * (CpuCFConv.cpp:160-178).
 *
 * Parity pin: tests/test_oracle_cfconv.py checks it against the SchNetPack golden outputs of the reference's own test
 * (TestCFConv.h:140-248 -> tests/golden/cfconv_water18.json) and against the compiled reference class (oracle/_ref).
 * Restated per centre atom over the FULL neighbour set (gather form); build with -DORACLE_REAL=double for the fp64 arbiter.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORACLE_REAL
#define ORACLE_REAL float
#endif
typedef ORACLE_REAL real;
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static real cf_displacement(const float* pos, const float* box, int a, int b, real d[3]) {
    for (int k = 0; k < 3; k++) d[k] = (real)pos[3 * b + k] - (real)pos[3 * a + k];
    if (box) {
        int tri = box[1] != 0 || box[2] != 0 || box[3] != 0 || box[5] != 0 || box[6] != 0 || box[7] != 0;
        real inv[3] = {(real)1 / box[0], (real)1 / box[4], (real)1 / box[8]};
        if (tri) {
            real s3 = (real)round(d[2] * inv[2]);
            d[0] -= s3 * box[6]; d[1] -= s3 * box[7]; d[2] -= s3 * box[8];
            real s2 = (real)round(d[1] * inv[1]);
            d[0] -= s2 * box[3]; d[1] -= s2 * box[4];
            real s1 = (real)round(d[0] * inv[0]);
            d[0] -= s1 * box[0];
        } else {
            d[0] -= (real)round(d[0] * inv[0]) * box[0];
            d[1] -= (real)round(d[1] * inv[1]) * box[4];
            d[2] -= (real)round(d[2] * inv[2]) * box[8];
        }
    }
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
}

/* filter f[c] and its r-derivative df[c] for one distance */
static void cf_filter(real r, int W, int G, real rc, real sigma, int activation, const float* w1, const float* b1, const float* w2,
                      const float* b2, real* f, real* df, real* y1, real* dy1, real* gs, real* dgs) {
    for (int g = 0; g < G; g++) {
        real mu = (real)g * rc / (G - 1);
        real x = (r - mu) / sigma;
        gs[g] = (real)exp(-0.5 * x * x);
        dgs[g] = -x * gs[g] / sigma;
    }
    for (int i = 0; i < W; i++) {
        real s = b1[i], ds = 0;
        for (int g = 0; g < G; g++) { s += gs[g] * w1[i * G + g]; ds += dgs[g] * w1[i * G + g]; }
        if (activation == 0) { real e = (real)exp(s); y1[i] = (real)log(0.5 * e + 0.5); dy1[i] = ds * e / (e + 1); }
        else { real t = (real)tanh(s); y1[i] = t; dy1[i] = ds * (1 - t * t); }
    }
    real fc = (real)(0.5 * cos(M_PI * r / rc) + 0.5), dfc = (real)(-(0.5 * M_PI / rc) * sin(M_PI * r / rc));
    for (int i = 0; i < W; i++) {
        real s = b2[i], ds = 0;
        for (int j = 0; j < W; j++) { s += y1[j] * w2[i * W + j]; ds += dy1[j] * w2[i * W + j]; }
        f[i] = fc * s;
        df[i] = dfc * s + fc * ds;
    }
}

/* output [n][W]; when output_grad != NULL also input_grad [n][W] and pos_grad [n][3] */
long long oracle_cfconv(int n, int W, int G, float cutoff, float sigma, int activation, const float* w1, const float* b1, const float* w2,
                        const float* b2, const float* pos, const float* box, const real* input, real* output, const real* output_grad,
                        real* input_grad, real* pos_grad) {
    real *f = malloc(sizeof(real) * W), *df = malloc(sizeof(real) * W), *y1 = malloc(sizeof(real) * W), *dy1 = malloc(sizeof(real) * W);
    real *gs = malloc(sizeof(real) * G), *dgs = malloc(sizeof(real) * G);
    const real rc2 = (real)cutoff * (real)cutoff;
    long long pairs = 0;
    memset(output, 0, sizeof(real) * (size_t)n * W);
    if (output_grad) { memset(input_grad, 0, sizeof(real) * (size_t)n * W); memset(pos_grad, 0, sizeof(real) * 3 * (size_t)n); }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            real d[3];
            real r2 = cf_displacement(pos, box, i, j, d);
            if (!(r2 < rc2)) continue;
            if (j > i) pairs++;
            real r = (real)sqrt(r2);
            cf_filter(r, W, G, cutoff, sigma, activation, w1, b1, w2, b2, f, df, y1, dy1, gs, dgs);
            for (int c = 0; c < W; c++) output[(size_t)i * W + c] += f[c] * input[(size_t)j * W + c];
            if (output_grad) {
                real w = 0;
                for (int c = 0; c < W; c++) {
                    input_grad[(size_t)i * W + c] += f[c] * output_grad[(size_t)j * W + c];
                    w += df[c] * (input[(size_t)j * W + c] * output_grad[(size_t)i * W + c] + input[(size_t)i * W + c] * output_grad[(size_t)j * W + c]);
                }
                w /= r;
                for (int k = 0; k < 3; k++) pos_grad[3 * i + k] -= w * d[k];
            }
        }
    free(f); free(df); free(y1); free(dy1); free(gs); free(dgs);
    return pairs;
}
