function [runHist, sigma] = dotsocp_gpu_level(var, opts, model, variant, method)
%% Run one level of the DOT-SOCP loop on the GPU (libdotsocp.so through mexDotSocpGPU).
% Drop-in body for socp/<variant>/algorithms/solver_*socp_*.m: same inputs, same side effects on the handle
% objects var/model and the same outputs (reference: solver_socp_inPALM.m:1, :328-357).
%   variant : 'dot2d' | 'wdot2d' | 'dot1d'          method : 'inPALM' | 'PALM' | 'acc-ADMM' | 'sGS-inPALM' | 'acc-sGS-ADMM'
% (ALG2 is inPALM with opts.tau = 1, exactly as in solver_dotsocp2d.m:133-137.)

P = struct();
P.variant = variant;
P.method  = method;
P.nt = model.nt;  P.nx = model.nx;
if strcmp(variant, 'dot1d'), P.ny = 1; else, P.ny = model.ny; end
P.maxit = opts.maxit;  P.tol = opts.tol;  P.sigma = opts.sigma;
if isfield(opts, 'tau'), P.tau = opts.tau; else, P.tau = 1; end
P.ifCheckStepByStep = double(opts.ifCheckStepByStep);
P.scaling = double(isfield(opts, 'scaling') && opts.scaling);
if isfield(opts, 'checkPrimDualFeas'), P.checkPrimDualFeas = double(opts.checkPrimDualFeas); else, P.checkPrimDualFeas = -1; end
if isfield(opts, 'time_limit'), P.time_limit = opts.time_limit; else, P.time_limit = NaN; end   % NaN = absent (3600 s)
if isfield(opts, 'restart'), P.restart = opts.restart; else, P.restart = 0; end
if isfield(opts, 'rho'),     P.rho = opts.rho;         else, P.rho = 0; end
if isfield(opts, 'theta'),   P.theta = opts.theta;     else, P.theta = 0; end
P.cScale = var.cScale;  P.dScale = var.dScale;  P.D = var.D;  P.E = var.E;
P.normc = model.normc;
if isprop(model, 'normd') && ~isempty(model.normd), P.normd = model.normd; else, P.normd = 0; end

% model.grad = D * [kron(Dt,I); kron(I,Dx,I); kron(I,Dy)] (initialize.m:35-39, solver_dotsocp2d.m:338): only the three
% magnitudes are needed on the device
lenT = (P.nt - 1) * P.nx * P.ny;
lenX = P.nt * (P.nx - 1) * P.ny;
P.grad_t = full(max(abs(model.grad(1, :))));
P.grad_x = full(max(abs(model.grad(lenT + 1, :))));
if P.ny > 1, P.grad_y = full(max(abs(model.grad(lenT + lenX + 1, :)))); else, P.grad_y = 0; end

% detach the iterates from the handle (:89-93): after this the local variables hold the ONLY reference to each array, so the
% in-place MEX write below cannot alias another MATLAB variable and no private copy is needed (the reference loop relies on
% exactly this; a forced copy would double the host footprint, 120 GB at 1024x1024x512).  The multilevel drivers create
% these arrays themselves (initialize.m / jump_nextLevel.m) and keep no second handle to them.
phi = var.phi;    var.phi   = [];
q = var.q;        var.q     = [];
z = var.z;        var.z     = [];
alpha = var.alpha; var.alpha = [];
beta = var.beta;  var.beta  = [];

if strcmp(variant, 'wdot2d'), weight = model.weight; else, weight = []; end

out = mexDotSocpGPU(phi, q, z, alpha, beta, model.c, weight, P);   % phi,q,z,alpha,beta are updated in place

%% output (solver_socp_inPALM.m:328-357)
names3 = struct('inPALM', 'Inexact Proximal ALM', 'PALM', 'Proximal ALM', 'accADMM', 'Accelerated ADMM', ...
                'sGSinPALM', 'Symmetric Gauss-seidel based inPALM', 'accsGSADMM', 'Accelerated symmetric Gauss-Seidel based ADMM');
var.name = names3.(strrep(strrep(method, '-', ''), 'ALG2', 'inPALM'));
var.phi = phi;  var.q = q;  var.z = z;  var.alpha = alpha;  var.beta = beta;   % alpha, beta already times sigma
switch method
    case 'PALM'
        tnames = {'Step_1_Q_Step', 'Step_2_1_FFT', 'Step_2_2_ProjSOC', 'Step_3_Q_Step', 'Step_4_Multiplier', 'KKT', 'Total_Time', 'Iters'};
        times = [out.times(1:7), out.iters];
    case 'acc-ADMM'
        tnames = {'Step_1_Q_Step', 'Step_2_Multiplier', 'Step_3_1_FFT', 'Step_3_2_ProjSOC', 'KKT', 'Interp', 'Total_Time', 'Iters'};
        times = [out.times(1:7), out.iters];
    case 'acc-sGS-ADMM'
        tnames = {'Step_1_1_sGS', 'Step_1_2_ProjSOC', 'Step_2_Multiplier', 'Step_3_Q_Step', 'Step_4_Interp', 'KKT', 'Total_Time', 'Iters'};
        times = [out.times(1:7), out.iters];
    case 'sGS-inPALM'
        tnames = {'Step_1_1_sGS', 'Step_1_2_ProjSOC', 'Step_2_Q_Step', 'Step_3_Multiplier', 'KKT', 'Total_Time', 'Iters'};
        times = [out.times(1:6), out.iters];
    otherwise
        tnames = {'Step_1_1_FFT', 'Step_1_2_ProjSOC', 'Step_2_Q_Step', 'Step_3_Multiplier', 'KKT', 'Total_Time', 'Iters'};
        times = [out.times(1:6), out.iters];
end
var.time = record_time(times, tnames);
var.cScale = out.cScale;  var.dScale = out.dScale;  var.D = out.D;  var.E = out.E;

n = out.len;
runHist = {};
kkt = out.kkt_t.';                 % the gateway fills rows contiguously (7 x maxit column-major)
runHist.kkt     = kkt(1:n, :);
runHist.time    = out.time(1:n);
runHist.iter    = out.iter(1:n);
runHist.pdGap   = out.pdGap(1:n);
runHist.priVal  = out.priVal(1:n);   % extras: objective values of :265-266
runHist.dualVal = out.dualVal(1:n);
runHist.len     = n;
sigma = out.sigma;

end
