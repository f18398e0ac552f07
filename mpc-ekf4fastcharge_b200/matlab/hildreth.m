function [DU, lambda, nexec] = hildreth(E, F, M, gamma, lambda0, maxIter)
% HILDRETH  drop-in for MPC-EKF4FastCharge/hildreth.m:1 -- same arguments and defaults
% (lambda0 = [] -> zeros, hildreth.m:18-22; maxIter = 100, hildreth.m:23-25); runs in libmek.
if nargin < 5, lambda0 = []; end
if nargin < 6, maxIter = 100; end
[DU, lambda, nexec] = mek_mex('hildreth', E.', F(:), M.', gamma(:), lambda0(:), maxIter);
nexec = double(nexec);
end
