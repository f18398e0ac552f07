function [Phi, G, aug] = predMat(A, B, C, D, Np, Nc)
% PREDMAT  drop-in for MPC-EKF4FastCharge/predMat.m:1 for the form the control loop uses:
% diagonal A, B = ones (enforced for every bank at initKF.m:74-88), one output row.
if ~isequal(diag(diag(A)), A) || any(B(:) ~= 1) || size(C,1) ~= 1 || size(A,1) ~= 6
  error('mek:predMat', 'libmek implements predMat for diagonal 6x6 A, B = ones, scalar output');
end
[PhiT, GT] = mek_mex('pred_mat', diag(A), C(:), D, Np, Nc);
Phi = PhiT.'; G = GT.';
m = 1; nx = 6;
aug = struct('A', [A, B; zeros(m,nx), eye(m)], 'B', [zeros(nx,m); eye(m)], 'C', [C, D], 'nx_aug', nx+m, 'm', m, 'q', 1);
end
