function [runHist, sigma] = solver_socp_accsGSADMM(var, opts, model)
%% GPU drop-in for socp/dot2d/algorithms/solver_socp_accsGSADMM.m (acc-ADMM with a red-black Gauss-Seidel phi-step)
% Place this file (with dotsocp_gpu_level.m and the built mexDotSocpGPU) in a copy of the reference's socp/ tree: the
% drivers add their own algorithms/ directory to the front of the path (solver_dotsocp2d.m:37-54), so the replacement
% must keep this file name.  Demos and drivers stay byte-identical.
[runHist, sigma] = dotsocp_gpu_level(var, opts, model, 'dot2d', 'acc-sGS-ADMM');
end
