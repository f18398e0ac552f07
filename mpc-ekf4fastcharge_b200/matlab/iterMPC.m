function [uk, mpcData] = iterMPC(xk, cellState, mpcData) %#ok<INUSL>
% ITERMPC  drop-in for MPC-EKF4FastCharge/iterMPC.m:1.  cellState.MPC is the struct from EKFmatsHandler,
% cellState.pop the population handle; the dual warm start (mpcData.lambda) lives in the device state.
[uk, phise, nexec, cost] = mek_mex('iter_mpc', cellState.pop, cellState.MPC.mats, mpcData.uk_1 .* ones(1, cellState.B));
mpcData.uk_1 = uk; mpcData.nexec = double(nexec); mpcData.phise = phise;       % iterMPC.m:69,78
if isfield(mpcData, 'k')                                                       % iterMPC.m:88-95 (rows = sample instants, columns = cells)
  k = mpcData.k;
  mpcData.cost.t(k, 1) = (k - 1) * mpcData.Ts;
  mpcData.cost.J_uncon(k, :) = cost(1, :); mpcData.cost.J_final(k, :) = cost(2, :);
  mpcData.cost.norm_DU(k, :) = cost(3, :); mpcData.cost.viol(k, :) = cost(4, :);
  mpcData.cost.nexec(k, :) = double(nexec);
end
end
