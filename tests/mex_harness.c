/* mex_harness.c - test infrastructure: drives ONE gateway of mex/gsmcal_mex.c through mexFunction() with arrays read from a
 * file and writes the returned arrays to another file, so a Python test can feed real captures to the gateway and compare
 * its outputs with the oracle.  Compiled together with gsmcal_mex.c against mex/stub/mex.h (no MATLAB/Octave here), in both
 * complex layouts (-DMX_HAS_INTERLEAVED_COMPLEX=0/1).
 *
 * file format (little endian): int32 count, int32 nlhs (input file only); per array: int32 class (6 double, 9 uint8,
 * 3 logical), int32 is_complex, int32 ndim, int64 dims[ndim], real part, imaginary part (always split in the FILE). */
#include <stdint.h>
#include "mex.h"

static size_t esize(mxClassID c) { return c == mxDOUBLE_CLASS ? 8 : 1; }

static mxArray *read_array(FILE *f) {
    int32_t cls, cplx, ndim; int64_t d64[3]; size_t dims[3] = {1, 1, 1};
    if (fread(&cls, 4, 1, f) != 1 || fread(&cplx, 4, 1, f) != 1 || fread(&ndim, 4, 1, f) != 1) exit(4);
    if (fread(d64, 8, (size_t)ndim, f) != (size_t)ndim) exit(4);
    for (int i = 0; i < ndim; ++i) dims[i] = (size_t)d64[i];
    mxArray *a = mxCreateNumericArray((size_t)ndim, dims, (mxClassID)cls, cplx ? mxCOMPLEX : mxREAL);
    size_t n = mxGetNumberOfElements(a), es = esize((mxClassID)cls);
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
    if (cplx) {
        double *tmp = (double *)malloc(2 * n * 8 + 8), *dst = (double *)a->re;
        if (fread(tmp, 8, 2 * n, f) != 2 * n) exit(4);
        for (size_t i = 0; i < n; ++i) { dst[2 * i] = tmp[i]; dst[2 * i + 1] = tmp[n + i]; }
        free(tmp);
        return a;
    }
#endif
    if (n && fread(a->re, es, n, f) != n) exit(4);
    if (cplx && n && fread(a->im, es, n, f) != n) exit(4);
    return a;
}

static void write_array(FILE *f, const mxArray *a) {
    int32_t cls = (int32_t)a->cls, cplx = a->is_complex, ndim = (int32_t)a->ndim; int64_t d64[3];
    for (int i = 0; i < ndim; ++i) d64[i] = (int64_t)a->dims[i];
    fwrite(&cls, 4, 1, f); fwrite(&cplx, 4, 1, f); fwrite(&ndim, 4, 1, f); fwrite(d64, 8, (size_t)ndim, f);
    size_t n = mxGetNumberOfElements(a), es = esize(a->cls);
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
    if (cplx) {
        const double *src = (const double *)a->re;
        for (int part = 0; part < 2; ++part) for (size_t i = 0; i < n; ++i) fwrite(&src[2 * i + part], 8, 1, f);
        return;
    }
#endif
    fwrite(a->re, es, n, f);
    if (cplx) fwrite(a->im, es, n, f);
}

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t nrhs, nlhs;
    if (fread(&nrhs, 4, 1, f) != 1 || fread(&nlhs, 4, 1, f) != 1) return 4;
    const mxArray *prhs[16]; mxArray *plhs[16] = {0};
    for (int i = 0; i < nrhs; ++i) prhs[i] = read_array(f);
    fclose(f);
    mexFunction(nlhs, plhs, nrhs, prhs);
    FILE *o = fopen(argv[2], "wb");
    int32_t n_out = 0;
    for (int i = 0; i < 16; ++i) if (plhs[i]) n_out = i + 1;
    fwrite(&n_out, 4, 1, o);
    for (int i = 0; i < n_out; ++i) write_array(o, plhs[i]);
    fclose(o);
    if (mex_at_exit_fn) mex_at_exit_fn();
    return 0;
}
