function [uk, mpcData] = iterMPC(xk, cellState, mpcData) %#ok<INUSL>
% ITERMPC  drop-in for MPC-EKF4FastCharge/iterMPC.m:1.  cellState.MPC is the struct from EKFmatsHandler,
% cellState.pop the population handle; the dual warm start (mpcData.lambda) lives in the device state.
[uk, phise, nexec] = mek_mex('iter_mpc', cellState.pop, cellState.MPC.mats, mpcData.uk_1 .* ones(1, cellState.B));
mpcData.uk_1 = uk; mpcData.nexec = double(nexec); mpcData.phise = phise;       % iterMPC.m:69,78
end
