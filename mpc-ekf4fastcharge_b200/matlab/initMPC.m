function mpcData = initMPC(SOC0, Np, Nc, targetSOC, opts)
% INITMPC  drop-in for MPC-EKF4FastCharge/initMPC.m:1 (same fields; limits may be 1xB rows).
if nargin < 5, opts = struct(); end
Np = max(1, round(Np)); Nc = max(1, min(round(Nc), Np));                       % initMPC.m:32-33
mpcData = struct('Np', Np, 'Nc', Nc, 'Ts', 1, 'ref', targetSOC, 'maxHild', 100, 'lambda', [], 'uk_1', 0, 'SOCk_1', 0, ...
                 'SOC0', SOC0, 'nexec', 0, 'k', 0);
if isfield(opts, 'Ts'), mpcData.Ts = opts.Ts; end
cons = [1 1 1]; if isfield(opts, 'constraints'), cons = opts.constraints; end
L = struct; if isfield(opts, 'limits'), L = opts.limits; end
L.u_min = -opts.ROM.Q * opts.Crate;                                            % initMPC.m:66-67
L.constraints = logical(cons(:)).';
mpcData.const = L;
end
