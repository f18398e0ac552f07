function ROM = mek_load_rom(bankfile)
% MEK_LOAD_ROM  replaces load(ROMfile,'ROM') (runMPC.m:4-5): opens a MEKBANK1 file in libmek.
ROM = struct('bank', mek_mex('bank_load', bankfile), 'nz', 30, 'Q', 29.86);
end
