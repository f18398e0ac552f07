% runMPC (libmek): MPC-EKF4FastCharge/runMPC.m with the k-loop (runMPC.m:83-112) replaced by ONE call
% into the fused CUDA kernel.  Same configuration literals as the reference (runMPC.m:8-50).
ROM = mek_load_rom('bank_nmc30_synth.bin');
TC = 25; Ts = 1; finalTime = 3000; time = (0:finalTime).'; N = numel(time); SOC0 = 10;
SigmaX0 = diag([ones(1,5) 2e6]); SigmaW = 1e2; SigmaV = 1e-3;
ekfData = initKF(SOC0, TC, SigmaX0, SigmaV, SigmaW, 'OutB', ROM);
Np = 5; Nc = 2; targetSOC = 95; constraints = [1 1 1]; Crate = 2;
limits = struct('u_max', 2, 'du_min', -50, 'du_max', 50, 'v_min', 3.4, 'v_max', 4.1, 'phise_min', 0.08, ...
                'z_max', targetSOC/100, 'z_tol', 0);
opts = struct('Ts', Ts, 'constraints', constraints, 'limits', limits, 'ROM', ROM, 'Crate', Crate, 'Nsim', finalTime);
mpcData = initMPC(SOC0, Np, Nc, targetSOC, opts);
initCfg = struct('SOC0', SOC0, 'warnOff', true, 'ekfData', ekfData, 'mpcData', mpcData);
[~, ~, cellState] = OB_step(0, TC, [], ROM, initCfg);
mek_mex('set_logging', cellState.pop, cellState.B, N);
mek_mex('run', cellState.pop, 0, N);                                           % runMPC.m:83-112 on the GPU
[u_store, voltage_store, SOC_store, phise_store, nexec] = mek_mex('read_logs', cellState.pop, 0, N, cellState.B);
u_store = u_store.'; voltage_store = voltage_store.'; phise_store = phise_store.';
plot4Figures(time, u_store, voltage_store, SOC_store, phise_store, mpcData);   % the reference's own plotting code
