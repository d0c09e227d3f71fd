/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
 *
 * CPU restatement of the ANI atomic-environment-vector (AEV) forward and backward pass, the
 * algorithm of the reference's `CpuANISymmetryFunctions` (reference paths relative to
 * /root/reference/src/ani/):
 *   - forward driver, scale factors ............ CpuANISymmetryFunctions.cpp:46-110
 *   - radial terms + angular neighbour set ..... CpuANISymmetryFunctions.cpp:112-151
 *   - angular terms ............................ CpuANISymmetryFunctions.cpp:153-194
 *   - backward driver .......................... CpuANISymmetryFunctions.cpp:196-226
 *   - radial backward .......................... CpuANISymmetryFunctions.cpp:228-263
 *   - angular backward ......................... CpuANISymmetryFunctions.cpp:265-353
 *   - minimum image, cutoff, angle helpers ..... CpuANISymmetryFunctions.cpp:355-439
 *
 * Parity pin: checked in tests/test_oracle_ani.py against (i) the TorchANI golden AEVs the
 * reference's own test holds (TestANISymmetryFunctions.h:111-252, transcribed to
 * tests/golden/ani_water18.json) and (ii) the reference CPU class itself compiled from
 * /root/reference into oracle/_ref (see oracle/Makefile).
 *
 * The file is written as a restatement, not a copy: the pair scan is organised as
 * "per centre atom, full neighbour row" and the backward is a per-centre gather; results agree
 * with the reference to fp32 round-off.  Build with -DORACLE_REAL=double for the fp64 arbiter.
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

typedef struct {
    int n_atoms, n_species, periodic, triclinic, torchani;
    real rcr, rca;
    real box[3][3], inv[3];
    const float* pos;
} ctx_t;

/* minimum image, CpuANISymmetryFunctions.cpp:355-379 (multiply by the reciprocal of the diagonal) */
static real displacement(const ctx_t* c, int a, int b, real d[3]) {
    for (int k = 0; k < 3; k++) d[k] = (real)c->pos[3 * b + k] - (real)c->pos[3 * a + k];
    if (c->periodic) {
        if (c->triclinic) {
            real s3 = (real)round(d[2] * c->inv[2]);
            d[0] -= s3 * c->box[2][0]; d[1] -= s3 * c->box[2][1]; d[2] -= s3 * c->box[2][2];
            real s2 = (real)round(d[1] * c->inv[1]);
            d[0] -= s2 * c->box[1][0]; d[1] -= s2 * c->box[1][1];
            real s1 = (real)round(d[0] * c->inv[0]);
            d[0] -= s1 * c->box[0][0];
        } else {
            for (int k = 0; k < 3; k++) d[k] -= (real)round(d[k] * c->inv[k]) * c->box[k][k];
        }
    }
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
}

/* cosine cutoff and its derivative, CpuANISymmetryFunctions.cpp:381-387 */
static real fcut(real r, real rc) { return (real)(0.5 * cos(M_PI * r / rc) + 0.5); }
static real dfcut(real r, real rc) { return (real)(-(0.5 * M_PI / rc) * sin(M_PI * r / rc)); }

/* angle between two displacement vectors, CpuANISymmetryFunctions.cpp:389-408 */
static real angle(const ctx_t* c, const real u[3], const real v[3], real ru, real rv) {
    real dot = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
    if (c->torchani) dot *= (real)0.95;
    real cosv = dot / (ru * rv);
    if (!c->torchani && (cosv > (real)0.99 || cosv < (real)-0.99)) {
        real x = u[1] * v[2] - u[2] * v[1], y = u[2] * v[0] - u[0] * v[2], z = u[0] * v[1] - u[1] * v[0];
        real a = (real)asin(sqrt(x * x + y * y + z * z) / (ru * rv));
        return cosv < 0 ? (real)(M_PI - a) : a;
    }
    return (real)acos(cosv);
}

/* d(angle)/d(u), d(angle)/d(v), CpuANISymmetryFunctions.cpp:410-433 */
static void angle_grads(const ctx_t* c, const real u[3], const real v[3], real ru, real rv, real gu[3], real gv[3]) {
    real dot = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
    real iu = 1 / ru, iv = 1 / rv, ip = iu * iv;
    real k;
    if (c->torchani) { real sd = (real)0.95 * dot * ip; k = (real)(-0.95 / sqrt(1 - sd * sd)); }
    else             { real sd = dot * ip;              k = (real)(-1 / sqrt(1 - sd * sd)); }
    for (int a = 0; a < 3; a++) {
        gu[a] = k * ip * (v[a] - dot * iu * iu * u[a]);
        gv[a] = k * ip * (u[a] - dot * iv * iv * v[a]);
    }
}

static int pair_index(int S, int s, int t) { /* CpuANISymmetryFunctions.cpp:39-43 */
    if (s > t) { int x = s; s = t; t = x; }
    return s * S - s * (s - 1) / 2 + (t - s);
}

static void init_ctx(ctx_t* c, int n_atoms, int n_species, float rcr, float rca, int torchani,
                     const float* pos, const float* box) {
    c->n_atoms = n_atoms; c->n_species = n_species; c->rcr = rcr; c->rca = rca; c->torchani = torchani;
    c->pos = pos; c->periodic = box != NULL; c->triclinic = 0;
    if (box) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            c->box[i][j] = box[3 * i + j];
            if (i != j && box[3 * i + j] != 0) c->triclinic = 1;   /* :59-64 */
        }
        /* the reference forms the reciprocal in fp32 (:52-54); keep that in the fp32 build */
        for (int i = 0; i < 3; i++) c->inv[i] = (real)1 / c->box[i][i];
    }
}

/*
 * radial_fn: n_radial x {eta, rs};  angular_fn: n_angular x {eta, rs, zeta, thetas}
 * radial  out: [n_atoms][n_species][n_radial]
 * angular out: [n_atoms][n_species(n_species+1)/2][n_angular]
 */
void oracle_ani_forward(int n_atoms, int n_species, float rcr, float rca, int torchani, const int* species,
                        int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                        const float* pos, const float* box, real* radial, real* angular) {
    ctx_t c; init_ctx(&c, n_atoms, n_species, rcr, rca, torchani, pos, box);
    const int n_pairs = n_species * (n_species + 1) / 2;
    const real rcr2 = c.rcr * c.rcr, rca2 = c.rca * c.rca;
    int* nb = (int*)malloc(sizeof(int) * (size_t)(n_atoms > 0 ? n_atoms : 1));
    for (int i = 0; i < n_atoms; i++) {
        real* ri = radial + (size_t)i * n_species * n_radial;
        real* ai = angular + (size_t)i * n_pairs * n_angular;
        memset(ri, 0, sizeof(real) * n_species * n_radial);
        memset(ai, 0, sizeof(real) * n_pairs * n_angular);
        int n_nb = 0;
        for (int j = 0; j < n_atoms; j++) {
            if (j == i) continue;
            real d[3];
            /* the reference evaluates pos[max]-pos[min]; r2 is symmetric so only the sign of d differs */
            real r2 = displacement(&c, i, j, d);
            if (!(r2 < rcr2)) continue;                        /* strict, :129 */
            if (r2 < rca2) nb[n_nb++] = j;                     /* nested test, :132-135 */
            real r = (real)sqrt(r2), fc = fcut(r, c.rcr);
            for (int k = 0; k < n_radial; k++) {
                real sh = r - radial_fn[2 * k + 1];
                ri[species[j] * n_radial + k] += fc * (real)exp(-radial_fn[2 * k] * sh * sh);
            }
        }
        if (torchani) for (int k = 0; k < n_species * n_radial; k++) ri[k] *= (real)0.25;   /* :99-103 */
        for (int a = 0; a < n_nb; a++) {
            real da[3]; real ra = (real)sqrt(displacement(&c, i, nb[a], da)); real fa = fcut(ra, c.rca);
            for (int b = a + 1; b < n_nb; b++) {
                real db[3]; real rb = (real)sqrt(displacement(&c, i, nb[b], db)); real fb = fcut(rb, c.rca);
                real rm = (real)0.5 * (ra + rb);
                real th = angle(&c, da, db, ra, rb);
                real* out = ai + pair_index(n_species, species[nb[a]], species[nb[b]]) * n_angular;
                for (int m = 0; m < n_angular; m++) {
                    const float* f = angular_fn + 4 * m;
                    real ct = (real)pow(1 + cos(th - f[3]), f[2]);
                    real sh = rm - f[1];
                    out[m] += fa * fb * ct * (real)exp(-f[0] * sh * sh);
                }
            }
        }
        for (int p = 0; p < n_pairs; p++)                       /* :104-109 */
            for (int m = 0; m < n_angular; m++) ai[p * n_angular + m] *= (real)pow(2, 1 - angular_fn[4 * m + 2]);
    }
    free(nb);
}

/* pos_grad: [n_atoms][3] = d(sum radial*radial_grad + angular*angular_grad)/d(pos) */
void oracle_ani_backward(int n_atoms, int n_species, float rcr, float rca, int torchani, const int* species,
                         int n_radial, const float* radial_fn, int n_angular, const float* angular_fn,
                         const float* pos, const float* box, const real* radial_grad, const real* angular_grad,
                         real* pos_grad) {
    ctx_t c; init_ctx(&c, n_atoms, n_species, rcr, rca, torchani, pos, box);
    const int n_pairs = n_species * (n_species + 1) / 2;
    const real rcr2 = c.rcr * c.rcr, rca2 = c.rca * c.rca;
    const real gscale = torchani ? (real)0.25 : (real)1;
    const int c2r = n_species * n_radial, c2a = n_pairs * n_angular;
    int* nb = (int*)malloc(sizeof(int) * (size_t)(n_atoms > 0 ? n_atoms : 1));
    memset(pos_grad, 0, sizeof(real) * 3 * (size_t)n_atoms);
    for (int i = 0; i < n_atoms; i++) {
        int n_nb = 0;
        for (int j = 0; j < n_atoms; j++) {
            if (j == i) continue;
            real d[3]; real r2 = displacement(&c, i, j, d);
            if (!(r2 < rcr2)) continue;
            if (r2 < rca2) nb[n_nb++] = j;
            if (j < i) continue;                                /* each undirected pair once, :233-234 */
            real r = (real)sqrt(r2), ir = 1 / r, fc = fcut(r, c.rcr), dfc = dfcut(r, c.rcr);
            for (int k = 0; k < n_radial; k++) {
                real eta = radial_fn[2 * k], sh = r - radial_fn[2 * k + 1];
                real e = (real)exp(-eta * sh * sh);
                real dv = dfc * e - fc * 2 * eta * sh * e;
                real g = radial_grad[(size_t)i * c2r + species[j] * n_radial + k] + radial_grad[(size_t)j * c2r + species[i] * n_radial + k];
                real s = gscale * g * dv * ir;
                for (int a = 0; a < 3; a++) { pos_grad[3 * i + a] -= s * d[a]; pos_grad[3 * j + a] += s * d[a]; }
            }
        }
        for (int a = 0; a < n_nb; a++) {
            int ja = nb[a];
            real da[3]; real ra = (real)sqrt(displacement(&c, i, ja, da)); real ira = 1 / ra;
            real fa = fcut(ra, c.rca), dfa = dfcut(ra, c.rca);
            for (int b = a + 1; b < n_nb; b++) {
                int jb = nb[b];
                real db[3]; real rb = (real)sqrt(displacement(&c, i, jb, db)); real irb = 1 / rb;
                real fb = fcut(rb, c.rca), dfb = dfcut(rb, c.rca);
                real rm = (real)0.5 * (ra + rb);
                real th = angle(&c, da, db, ra, rb);
                real ga[3], gb[3]; angle_grads(&c, da, db, ra, rb, ga, gb);
                const real* g = angular_grad + (size_t)i * c2a + pair_index(n_species, species[ja], species[jb]) * n_angular;
                for (int m = 0; m < n_angular; m++) {
                    const float* f = angular_fn + 4 * m;
                    real base = (real)(1 + cos(th - f[3]));
                    real ct = (real)pow(base, f[2]);
                    real sh = rm - f[1];
                    real e = (real)exp(-f[0] * sh * sh);
                    real de = -f[0] * sh * e;
                    real zs = (real)pow(2, 1 - f[2]) * g[m];
                    real wa = zs * (dfa * fb * ct * e + fa * fb * ct * de) * ira;
                    real wb = zs * (fa * dfb * ct * e + fa * fb * ct * de) * irb;
                    real wt = zs * fa * fb * e * (real)(-f[2] * pow(base, f[2] - 1) * sin(th - f[3]));
                    for (int k = 0; k < 3; k++) {
                        real xa = wa * da[k] + wt * ga[k], xb = wb * db[k] + wt * gb[k];
                        pos_grad[3 * ja + k] += xa; pos_grad[3 * jb + k] += xb; pos_grad[3 * i + k] -= xa + xb;
                    }
                }
            }
        }
    }
    free(nb);
}
