%% Build the MEX gateway against libdotsocp.so (run inside MATLAB on the GPU machine, from the repository root).
%   1. make -C dotsocp_b200/csrc            (nvcc, sm_100a)  -> dotsocp_b200/libdotsocp.so
%   2. this script                                           -> matlab/mex/mexDotSocpGPU.mexa64
%   3. copy matlab/socp/<variant>/algorithms/*.m + matlab/socp/dotsocp_gpu_level.m + the MEX file over a copy of the
%      reference's socp/ tree (see INTEGRATION.md); LD_LIBRARY_PATH must contain dotsocp_b200/.
root = fileparts(fileparts(mfilename('fullpath')));
mex('-R2018a', fullfile(root, 'matlab', 'mex', 'mexDotSocpGPU.cpp'), ['-I' fullfile(root, 'include')], ...
    ['-L' fullfile(root, 'dotsocp_b200')], '-ldotsocp', '-outdir', fullfile(root, 'matlab', 'mex'));
