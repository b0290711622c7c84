/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): 80-bit extended-precision block-tridiagonal
 * products for the arbiter of oracle/solvers.py:shift_invert_extended.
 *
 * The reference applies OP = (A - sigma B)^-1 B with zgbmv + zgbtrs
 * (src/solvers/arnoldi/smod_arpack_shift_invert.f08:98-104).  The arbiter makes every
 * application forward-accurate by iterative refinement; the residual b - (A - sigma B) y and the
 * product B x are formed here in x87 long double (64-bit significand) from the *double* entries of
 * A and B, i.e. for the exact pencil the reference and the device both start from.
 *
 * Layout: blocks[b][w][i][j] (w = 0 sub, 1 diag, 2 super; row-major d x d), complex128 interleaved;
 * vectors are C `long double _Complex` (numpy clongdouble on x86-64, 32 bytes per entry).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC oracle/extprec.c -o oracle/_build/libextprec.so
 */
#include <stddef.h>

typedef struct { long double re, im; } cld;

/* y = alpha_sign * (A - sigma B) x + (z ? z : 0), with A - sigma B formed in long double per entry.
 * B == NULL: the matrix is A alone.  sign = +1 gives the product, sign = -1 with z = b the residual. */
void bt_gemv_ld(int G, int d, const double *A, const double *B, const double *sigma_ri, int sign,
                const cld *x, const cld *z, cld *y)
{
    const long double sr = sigma_ri ? sigma_ri[0] : 0.0L, si = sigma_ri ? sigma_ri[1] : 0.0L;
    const size_t bs = (size_t)d * d;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < G; ++b) {
        for (int i = 0; i < d; ++i) {
            long double ar = 0.0L, ai = 0.0L;
            for (int w = 0; w < 3; ++w) {
                const int bc = b + w - 1;
                if (bc < 0 || bc >= G) continue;
                const double *pa = A + 2 * (((size_t)b * 3 + w) * bs + (size_t)i * d);
                const double *pb = B ? B + 2 * (((size_t)b * 3 + w) * bs + (size_t)i * d) : NULL;
                const cld *xv = x + (size_t)bc * d;
                for (int j = 0; j < d; ++j) {
                    long double mr = pa[2 * j], mi = pa[2 * j + 1];
                    if (pb) {
                        const long double br = pb[2 * j], bi = pb[2 * j + 1];
                        mr -= sr * br - si * bi;
                        mi -= sr * bi + si * br;
                    }
                    ar += mr * xv[j].re - mi * xv[j].im;
                    ai += mr * xv[j].im + mi * xv[j].re;
                }
            }
            const size_t row = (size_t)b * d + i;
            cld out;
            out.re = sign * ar;
            out.im = sign * ai;
            if (z) { out.re += z[row].re; out.im += z[row].im; }
            y[row] = out;
        }
    }
}
