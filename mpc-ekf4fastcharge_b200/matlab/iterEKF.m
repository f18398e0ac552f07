function [zk, boundzk, ekfData, Xind] = iterEKF(vk, ik, Tk, ekfData)
% ITEREKF  drop-in for UTILITY/iterEKF.m:30 (output-blend mode).  ekfData.pop must hold the population
% handle (set it from cellState.pop after OB_step's first call).  Xind carries the linearisation that
% EKFmatsHandler consumes (EKFmatsHandler.m:26-114 runs in the same kernel).
B = numel(vk);
[zk, boundzk, mats] = mek_mex('iter_ekf', ekfData.pop, vk(:).', ik .* ones(1, B), ekfData.nz);
ekfData.priorI = ik;                                                           % iterEKF.m:210
Xind = struct('mats', mats, 'theT', mats(2,:), 'theZ', mats(3,:), 'gamma', NaN(4, B));
end
