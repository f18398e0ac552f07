function kfData = initKF(SOC0, T0, SigmaX0, SigmaV, SigmaW, blend, ROMs)
% INITKF  drop-in for UTILITY/initKF.m:30.  ROMs is the struct returned by mek_load_rom (it carries
% the libmek bank handle).  The device population is created lazily by OB_step's first call, once
% initMPC has supplied the horizon and limits (runMPC.m:20,52,74); SOC0/T0/SigmaX0 may be 1xB / 6x6xB.
switch upper(char(blend))                                                             % initKF.m:44-53
  case {'OUTB','OB'}, method = 'OB';
  case {'MDLB','MB'}, method = 'MB';   % model blend: served by iterEKF one call at a time (the fused loop is output-blend)
  otherwise
    warning('"blend" input to initKF not recognized. Switching to output blending (OutB)'); method = 'OB';
end
if size(SigmaX0,1) ~= 6 || size(SigmaX0,2) ~= 6
  error('Input argument SigmaX0 has wrong dimension. Should be %d by %d', 6, 6);   % initKF.m:66-69
end
kfData = struct('SOC0', SOC0(:).'/100, 'T0', T0(:).', 'SigmaX0', SigmaX0, 'SigmaV', SigmaV, 'SigmaW', SigmaW, ...
                'method', method, 'nz', ROMs.nz, 'n', 5, 'bank', ROMs.bank, 'pop', uint64(0), 'priorI', 0);
end
