/* TEST INFRASTRUCTURE ONLY (oracle/) -- never linked into the product.
 *
 * Minimal stand-in for MATLAB's libmx.so / libmex.so so that the reference's
 * pre-built, source-less MEX kernels (socp/<variant>/utils/mex*.mexa64) can be
 * dlopen()ed from Python (ctypes) and driven through their own
 * mexFunction(nlhs, plhs, nrhs, prhs) entry point.
 *
 * The binaries import only these accessors (nm -D):
 *   2-D kernels : mxGetPr, mxGetScalar, mxGetM, mxGetN
 *   1-D kernels : + mxIsDouble, mxGetNumberOfElements, mexErrMsgIdAndTxt
 * so a fake mxArray carrying {data pointer, rows, cols} is sufficient.
 *
 * Build (see oracle/Makefile): the SONAME must be libmx.so / libmex.so, otherwise the
 * dynamic loader cannot satisfy the NEEDED entries of the .mexa64 files from an
 * already-loaded library.
 */
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double *pr;
    size_t  m, n;
} mxArray;

/* last error raised through mexErrMsgIdAndTxt (the 1-D kernels validate their inputs) */
static char g_last_err_id[128];
static char g_last_err_msg[512];
static int  g_err_count = 0;

double *mxGetPr(const mxArray *a) { return a->pr; }
double *mxGetDoubles(const mxArray *a) { return a->pr; }
double  mxGetScalar(const mxArray *a) { return a->pr[0]; }
size_t  mxGetM(const mxArray *a) { return a->m; }
size_t  mxGetN(const mxArray *a) { return a->n; }
unsigned char mxIsDouble(const mxArray *a) { (void)a; return 1; }
unsigned char mxIsComplex(const mxArray *a) { (void)a; return 0; }
size_t  mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    va_list ap;
    strncpy(g_last_err_id, id ? id : "", sizeof g_last_err_id - 1);
    va_start(ap, fmt);
    vsnprintf(g_last_err_msg, sizeof g_last_err_msg, fmt ? fmt : "", ap);
    va_end(ap);
    g_err_count++;
    /* MATLAB would longjmp out of the MEX function here.  We cannot unwind a foreign
     * frame portably; callers in oracle/refmex.py validate arguments BEFORE calling so this
     * path is only reached by the deliberate error-behaviour probes, which run in a
     * forked child process. */
    fprintf(stderr, "mexErrMsgIdAndTxt: %s: %s\n", g_last_err_id, g_last_err_msg);
    fflush(stderr);
    _Exit(86);
}

void mexErrMsgTxt(const char *msg)
{
    mexErrMsgIdAndTxt("mex:error", "%s", msg);
}

int         mxshim_err_count(void) { return g_err_count; }
const char *mxshim_last_err_id(void) { return g_last_err_id; }
const char *mxshim_last_err_msg(void) { return g_last_err_msg; }
