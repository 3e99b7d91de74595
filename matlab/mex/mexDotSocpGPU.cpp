// mexDotSocpGPU.cpp -- thin MEX gateway from MATLAB to libdotsocp.so (C ABI in include/dotsocp.h).
//
//   out = mexDotSocpGPU(phi, q, z, alpha, beta, c, weight, P)
//
// phi, q, z, alpha, beta : the iterates of var (double; z/beta are L x 10, or L x 6 for the 1-D variant).  They are
//                          MODIFIED IN PLACE, exactly like the reference's own MEX kernels write into prhs[0]
//                          ("No output argument is required. pos 1 will be modified in-place", mexBFd1d.mexa64), so the
//                          .m wrapper must own the only reference (the reference loop detaches them from the handle
//                          for the same reason, solver_socp_inPALM.m:89-93).
// c                      : model.c (N x 1, already scaled);   weight : model.weight (Q x 1) or [] (unweighted)
// P                      : struct of scalars -- variant ('dot2d'|'wdot2d'|'dot1d'), method ('inPALM'|'PALM'|'acc-ADMM'|'sGS-inPALM'|'acc-sGS-ADMM'),
//                          nt,nx,ny, maxit, tol, tau, sigma, ifCheckStepByStep, scaling, checkPrimDualFeas (-1 = absent),
//                          time_limit, restart, rho, theta, cScale, dScale, D, E, normc, normd, grad_t, grad_x, grad_y
// out                    : struct with kkt (len x 7), time, iter, pdGap, priVal, dualVal (len x 1), len, iters, sigma,
//                          cScale, dScale, D, E, times (1 x 8), gpu_launches
//
// Replaces the body of socp/<variant>/algorithms/solver_*socp_*.m (the loops at solver_socp_inPALM.m:136-325 etc.).
// Errors are raised with mexErrMsgIdAndTxt("dotsocp:<code>", ...) (the 1-D reference kernels use the same mechanism,
// ids mexBFd:invalidNumInputs / invalidInput); there is no CPU fallback.
#include "mex.h"

#include <cstring>
#include <string>
#include <vector>

#include "dotsocp.h"

static double field(const mxArray* s, const char* name, double dflt, bool required)
{
    const mxArray* f = mxGetField(s, 0, name);
    if (!f || mxIsEmpty(f)) {
        if (required) mexErrMsgIdAndTxt("dotsocp:invalidInput", "P.%s is required", name);
        return dflt;
    }
    return mxGetScalar(f);
}

static std::string sfield(const mxArray* s, const char* name)
{
    const mxArray* f = mxGetField(s, 0, name);
    if (!f || !mxIsChar(f)) mexErrMsgIdAndTxt("dotsocp:invalidInput", "P.%s must be a char vector", name);
    char* c = mxArrayToString(f);
    std::string r(c ? c : "");
    mxFree(c);
    return r;
}

static double* dbl(const mxArray* a, size_t expect, const char* what)
{
    if (!a || !mxIsDouble(a) || mxIsComplex(a) || mxIsSparse(a))
        mexErrMsgIdAndTxt("dotsocp:invalidInput", "%s must be a full real double array", what);
    if (mxGetNumberOfElements(a) != expect)
        mexErrMsgIdAndTxt("dotsocp:invalidInput", "%s has %zu elements, expected %zu", what, mxGetNumberOfElements(a), expect);
    return mxGetPr(a);
}

// libdotsocp keeps the device session of the last (variant, grid) alive between calls; keep this MEX file (and with it the
// library's state) loaded for the life of the MATLAB process and free the device memory when MATLAB unloads it
static void at_exit() { dotsocp_release_cached(); }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    static bool locked = false;
    if (!locked) { mexLock(); mexAtExit(at_exit); locked = true; }
    if (nrhs != 8) mexErrMsgIdAndTxt("dotsocp:invalidNumInputs", "8 inputs required: phi,q,z,alpha,beta,c,weight,P");
    if (nlhs > 1) mexErrMsgIdAndTxt("dotsocp:invalidNumOutputs", "at most one output");
    const mxArray* P = prhs[7];
    if (!mxIsStruct(P)) mexErrMsgIdAndTxt("dotsocp:invalidInput", "P must be a struct");

    dotsocp_level_opts o;
    std::memset(&o, 0, sizeof o);
    const std::string variant = sfield(P, "variant"), method = sfield(P, "method");
    if (variant == "dot2d") o.variant = DOTSOCP_VARIANT_DOT2D;
    else if (variant == "wdot2d") o.variant = DOTSOCP_VARIANT_WDOT2D;
    else if (variant == "dot1d") o.variant = DOTSOCP_VARIANT_DOT1D;
    else mexErrMsgIdAndTxt("dotsocp:invalidInput", "unknown variant '%s'", variant.c_str());
    if (method == "inPALM" || method == "ALG2") o.method = DOTSOCP_METHOD_INPALM;
    else if (method == "PALM") o.method = DOTSOCP_METHOD_PALM;
    else if (method == "acc-ADMM") o.method = DOTSOCP_METHOD_ACCADMM;
    else if (method == "sGS-inPALM") o.method = DOTSOCP_METHOD_SGSINPALM;
    else if (method == "acc-sGS-ADMM") o.method = DOTSOCP_METHOD_ACCSGSADMM;
    else mexErrMsgIdAndTxt("dotsocp:invalidInput", "unknown method '%s'", method.c_str());
    o.nt = (int)field(P, "nt", 0, true);
    o.nx = (int)field(P, "nx", 0, true);
    o.ny = o.variant == DOTSOCP_VARIANT_DOT1D ? 1 : (int)field(P, "ny", 0, true);
    o.maxit = (int)field(P, "maxit", 0, true);
    o.ifCheckStepByStep = field(P, "ifCheckStepByStep", 0, false) != 0;
    o.scaling = field(P, "scaling", 0, false) != 0;
    o.checkPrimDualFeas = (int)field(P, "checkPrimDualFeas", -1, false);
    o.restart = (int)field(P, "restart", 0, false);
    o.tau = field(P, "tau", 1.0, false);
    o.sigma = field(P, "sigma", 0, true);
    o.tol = field(P, "tol", 0, true);
    o.time_limit = field(P, "time_limit", mxGetNaN(), false);   // NaN = absent (3600 s); <= 0 = budget already spent
    o.rho = field(P, "rho", 0, false);
    o.theta = field(P, "theta", 0, false);
    o.cScale = field(P, "cScale", 0, true); o.dScale = field(P, "dScale", 0, true);
    o.D = field(P, "D", 0, true); o.E = field(P, "E", 0, true);
    o.normc = field(P, "normc", 0, true); o.normd = field(P, "normd", 0, false);
    o.grad_t = field(P, "grad_t", 0, true); o.grad_x = field(P, "grad_x", 0, true); o.grad_y = field(P, "grad_y", 0, false);
    if (o.nt < 2 || o.nx < 2 || o.ny < 1 || o.maxit < 1) mexErrMsgIdAndTxt("dotsocp:invalidInput", "bad grid / maxit");

    const size_t N = (size_t)o.nt * o.nx * o.ny, L = (size_t)(o.nt - 1) * o.nx * o.ny;
    const size_t Q = L + (size_t)o.nt * (o.nx - 1) * o.ny + (size_t)o.nt * o.nx * (o.ny - 1);
    const size_t ncol = o.variant == DOTSOCP_VARIANT_DOT1D ? 6 : 10;
    double* phi = dbl(prhs[0], N, "phi");
    double* q = dbl(prhs[1], Q, "q");
    double* z = dbl(prhs[2], L * ncol, "z");
    double* alpha = dbl(prhs[3], Q, "alpha");
    double* beta = dbl(prhs[4], L * ncol, "beta");
    const double* c = dbl(prhs[5], N, "c");
    const double* weight = nullptr;
    if (o.variant == DOTSOCP_VARIANT_WDOT2D) weight = dbl(prhs[6], Q, "weight");

    const size_t cap = (size_t)o.maxit;
    mxArray* kkt_t = mxCreateDoubleMatrix(7, cap, mxREAL);   // row-major cap x 7 == column-major 7 x cap; the wrapper transposes
    mxArray* tim = mxCreateDoubleMatrix(cap, 1, mxREAL);
    mxArray* itr = mxCreateDoubleMatrix(cap, 1, mxREAL);
    mxArray* gap = mxCreateDoubleMatrix(cap, 1, mxREAL);
    mxArray* pri = mxCreateDoubleMatrix(cap, 1, mxREAL);
    mxArray* dua = mxCreateDoubleMatrix(cap, 1, mxREAL);
    dotsocp_hist h;
    h.cap = (int)cap;
    h.kkt = mxGetPr(kkt_t); h.time = mxGetPr(tim); h.iter = mxGetPr(itr); h.pdGap = mxGetPr(gap);
    h.priVal = mxGetPr(pri); h.dualVal = mxGetPr(dua);
    dotsocp_level_result r;
    const int rc = dotsocp_solve_level(&o, phi, q, z, alpha, beta, c, weight, &h, &r);
    if (rc != DOTSOCP_OK) mexErrMsgIdAndTxt("dotsocp:failed", "libdotsocp error %d: %s", rc, dotsocp_last_error());

    const char* names[] = {"kkt_t", "time", "iter", "pdGap", "priVal", "dualVal", "len", "iters", "sigma",
                           "cScale", "dScale", "D", "E", "times", "gpu_launches"};
    mxArray* out = mxCreateStructMatrix(1, 1, 15, names);
    mxSetField(out, 0, "kkt_t", kkt_t); mxSetField(out, 0, "time", tim); mxSetField(out, 0, "iter", itr);
    mxSetField(out, 0, "pdGap", gap); mxSetField(out, 0, "priVal", pri); mxSetField(out, 0, "dualVal", dua);
    mxSetField(out, 0, "len", mxCreateDoubleScalar(r.hist_len));
    mxSetField(out, 0, "iters", mxCreateDoubleScalar(r.iters));
    mxSetField(out, 0, "sigma", mxCreateDoubleScalar(r.sigma));
    mxSetField(out, 0, "cScale", mxCreateDoubleScalar(r.cScale));
    mxSetField(out, 0, "dScale", mxCreateDoubleScalar(r.dScale));
    mxSetField(out, 0, "D", mxCreateDoubleScalar(r.D));
    mxSetField(out, 0, "E", mxCreateDoubleScalar(r.E));
    mxArray* tm = mxCreateDoubleMatrix(1, DOTSOCP_NTIMES, mxREAL);
    std::memcpy(mxGetPr(tm), r.times, sizeof r.times);
    mxSetField(out, 0, "times", tm);
    mxSetField(out, 0, "gpu_launches", mxCreateDoubleScalar(r.gpu_launches));
    plhs[0] = out;
}
