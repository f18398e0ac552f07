function lut = mek_sample_celldata(ROM, n, TK)
% MEK_SAMPLE_CELLDATA  tabulate the function handles of ROM.cellData.function for libmek.
%   lut = mek_sample_celldata(ROM)         2001 points over theta = 0..1 at 298.15 K
%   lut = mek_sample_celldata(ROM, n, TK)
% The reference evaluates ROM.cellData.function.{neg,pos}.{Uocp,dUocp,k0,Rf}(theta,T) inside the loop
% (OB_step.m:231-232,329-330,337-340; iterEKF.m:362-363,392-393,404-407,441-442,463-464,495-496;
% EKFmatsHandler.m:60-61,68-69,84-85,96).  Function handles cannot cross the C ABI: this samples them
% on a uniform stoichiometry grid; libmek interpolates linearly (interp1 'linear', clamped to the grid).
%   lut.theta [1 x n], lut.tables [8 x n] in the order of include/mek.h (mek_bank_set_lut):
%   Uocp_neg dUocp_neg Uocp_pos dUocp_pos k0_neg k0_pos Rf_neg Rf_pos
% k0 and Rf are sampled at TK, which must be the reference temperature P_TREF of the parameter vector
% (the Arrhenius factors of the parameter vector are applied on top).
% Use:  rom = mek_load_rom(ROMfile, params);  mek_mex('bank_set_lut', rom.bank, lut.theta(1), lut.theta(end), lut.tables);
% or save('celldata_lut.mat','-struct','lut') and pass the file to mek/bank.py::from_mat(..., lut_mat=...).
if nargin < 2 || isempty(n),  n = 2001;   end
if nargin < 3 || isempty(TK), TK = 298.15; end
fn = ROM.cellData.function;
th = linspace(0, 1, n);
T = zeros(8, n);
for i = 1:n
  T(1, i) = fn.neg.Uocp(th(i), TK);  T(2, i) = fn.neg.dUocp(th(i), TK);
  T(3, i) = fn.pos.Uocp(th(i), TK);  T(4, i) = fn.pos.dUocp(th(i), TK);
  T(5, i) = fn.neg.k0(th(i), TK);    T(6, i) = fn.pos.k0(th(i), TK);
  T(7, i) = fn.neg.Rf(th(i), TK);    T(8, i) = fn.pos.Rf(th(i), TK);
end
lut = struct('theta', th, 'tables', T);
end
