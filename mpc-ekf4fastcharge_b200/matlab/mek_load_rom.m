function ROM = mek_load_rom(bankfile, lut)
% MEK_LOAD_ROM  replaces load(ROMfile,'ROM') (runMPC.m:4-5): opens a MEKBANK1 file in libmek.
%   ROM = mek_load_rom(bankfile)        closed-form cell functions of the bank's parameter vector (or the tables in the file)
%   ROM = mek_load_rom(bankfile, lut)   lut = mek_sample_celldata(realROM): the real cell's function handles, tabulated
ROM = struct('bank', mek_mex('bank_load', bankfile), 'nz', 30, 'Q', 29.86);
if nargin > 1 && ~isempty(lut)
  mek_mex('bank_set_lut', ROM.bank, lut.theta(1), lut.theta(end), lut.tables);
end
end
