function [runHist, sigma] = solver_socp_sGSinPALM(var, opts, model)
%% GPU drop-in for socp/dot2d/algorithms/solver_socp_sGSinPALM.m (phi-step = one red-black symmetric Gauss-Seidel sweep)
% Place this file (with dotsocp_gpu_level.m and the built mexDotSocpGPU) in a copy of the reference's socp/ tree: the
% drivers add their own algorithms/ directory to the front of the path (solver_dotsocp2d.m:37-54), so the replacement
% must keep this file name.  Demos and drivers stay byte-identical.
[runHist, sigma] = dotsocp_gpu_level(var, opts, model, 'dot2d', 'sGS-inPALM');
end
