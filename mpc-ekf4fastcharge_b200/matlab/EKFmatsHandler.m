function [MPC, xhat] = EKFmatsHandler(ekfData, Xind, zk, Tk) %#ok<INUSD>
% EKFMATSHANDLER  drop-in for MPC-EKF4FastCharge/EKFmatsHandler.m:1 -- unpacks what mek_pop_iter_ekf
% already computed on the device (fields as EKFmatsHandler.m:101-114, B columns for B cells).
m = Xind.mats; B = size(m, 2);
xhat = m(4:9, :);
MPC = struct('a', m(10:14, :), 'Csoc', [], 'Dsoc', 0, 'Cv', [m(15:19, :); zeros(1, B)], 'Dv', m(20, :), ...
             'Cphi', [m(21:25, :); zeros(1, B)], 'Dphi', m(26, :), 'bv', m(27, :), 'bphi', m(28, :), ...
             'iT', m(2, :), 'iZ', m(3, :), 'mats', m);
if B == 1
  MPC.A = diag([MPC.a; 1]); MPC.B = ones(6, 1); MPC.Cv = MPC.Cv.'; MPC.Cphi = MPC.Cphi.';
end
end
