function [Vcell, obs, cellState] = OB_step(Iapp, Tc, cellState, ROM, initCfg)
% OB_STEP  drop-in for MPC-EKF4FastCharge/OB_step.m:1.  On the first call (cellState = []) initCfg must
% also carry .ekfData and .mpcData (from initKF / initMPC): the three reference structs share ONE
% device state block, created here (mek_pop_create + mek_pop_init = initKF + initMPC + OB_step.m:39-72).
if nargin < 3 || isempty(cellState)
  if nargin < 5 || ~isfield(initCfg, 'SOC0'), error('First call requires initCfg.SOC0 (in %).'); end   % OB_step.m:40-42
  kf = initCfg.ekfData; mp = initCfg.mpcData; c = mp.const; B = numel(initCfg.SOC0);
  term = isfield(c, 'z_target_terminal') && c.z_target_terminal;                                       % constraintsMPC.m:104
  cfg = [mp.Np mp.Nc mp.maxHild double(c.constraints) 1 mp.ref kf.SigmaV kf.SigmaW double(strcmp(kf.method, 'MB')) double(term) c.z_tol];
  pop = mek_mex('pop_create', ROM.bank, cfg, B, 0);
  one = ones(1, B);
  lim = [c.u_max.*one; c.u_min.*one; c.du_max.*one; c.du_min.*one; c.v_max.*one; c.phise_min.*one; (c.z_max + c.z_tol).*one];
  sx = kf.SigmaX0; if size(sx, 3) == 1, sx = repmat(sx, 1, 1, B); end
  mek_mex('pop_init', pop, initCfg.SOC0(:).', kf.SOC0*100, Tc.*one, sx, lim);
  cellState = struct('initialized', true, 'pop', pop, 'B', B, 'nz', ROM.nz, 'MPC', []);
  Vcell = NaN(1, B); obs = struct();      % the reference's first call also steps with I = 0 (Q20): done inside pop_init
  return
end
Vcell = mek_mex('ob_step', cellState.pop, Iapp .* ones(1, cellState.B));
obs = struct('Vcell', Vcell);
end
