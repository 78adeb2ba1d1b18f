% make_golden_matlab.m -- golden vectors for the MATLAB-ONLY stages of the VBHEM E-step / EM iteration, produced by
% the reference toolbox's own files (Octave >= 6 or MATLAB):
%
%   pre-pass                    vbhem_h3m_c_step_fc.m:118-165,180-191
%   responsibilities z_hat, Nj  vbhem_h3m_c_step_fc.m:271-283
%   statistics reduction        vbhem_compute_Statistics.m:33-82
%   lower bound                 vbhemh3m_lb.m:64-186
%   M-step                      vbhem_mstep_component.m:42-70
%
% Usage (from the repository root, with <reference> = a checkout of the toolbox):
%   octave --eval "addpath(genpath('<reference>/src')); run('tests/golden/make_golden_matlab.m')"
% It reads tests/golden/matlab_stage_inputs.mat (tests/golden/make_matlab_inputs.py), runs ONE EM run of the
% reference (vbhem_h3m_c_step_fc with the MEX switched off, so that the MATLAB implementation of the per-pair part,
% vbhem_hmm_bwd_fwd_fast.m, is the one exercised) and writes tests/golden/matlab_stage_golden.mat:
%   LogLs [1 x iters]            the trace of lower bounds (h3m_r.LogLs)
%   alpha, eta, epsilon, m, W, v, lambda   the centre hyper-parameters after the run (layouts of the inputs)
%   Z [N x K], L_elbo [N x K], Nj [1 x K]  of the last E-step
% tests/test_matlab_golden.py then compares the CUDA path (vbhem_em_run) and the numpy restatement with it.
here = fileparts(mfilename('fullpath'));
in = load(fullfile(here, 'matlab_stage_inputs.mat'));

N = size(in.b_prior, 1);  d = size(in.b_mu, 3);
K = size(in.r_eta, 1);    S = size(in.r_eta, 2);

% ---- base mixture h3m_b (hmms_to_h3m_hem.m:19-140)
h3m_b = struct();
h3m_b.K = N;
h3m_b.omega = ones(1, N) / N;
h3m_b.hmm = cell(1, N);
for i = 1:N
  s = in.b_Sb(i);
  hmm = struct();
  hmm.prior = reshape(in.b_prior(i, 1:s), [s 1]);
  hmm.A     = reshape(in.b_A(i, 1:s, 1:s), [s s]);            % A(from, to)
  hmm.emit  = cell(1, s);
  for k = 1:s
    e = struct('type', 'gmm', 'nin', d, 'ncentres', 1, 'priors', 1, 'covar_type', 'full');
    e.centres = reshape(in.b_mu(i, k, :), [1 d]);
    e.covars  = reshape(in.b_cov(i, k, :, :), [d d]);
    hmm.emit{k} = e;
  end
  h3m_b.hmm{i} = hmm;
end

% ---- initial centres h3m_r (the fields vbhemhmm_init.m:76-99 sets)
h3m_r = struct();
h3m_r.K = K;  h3m_r.S = S;
h3m_r.alpha = reshape(in.r_alpha, [1 K]);
h3m_r.LogL = -inf;
h3m_r.hmm = cell(1, K);
for j = 1:K
  hmm = struct();
  hmm.eta     = reshape(in.r_eta(j, :), [S 1]);
  hmm.epsilon = reshape(in.r_epsilon(j, :, :), [S S]);        % epsilon(from, to)
  hmm.emit    = cell(1, S);
  for k = 1:S
    e = struct();
    e.m      = reshape(in.r_m(j, k, :), [1 d]);
    e.W      = reshape(in.r_W(j, k, :, :), [d d]);
    e.v      = in.r_v(j, k);
    e.lambda = in.r_lam(j, k);
    hmm.emit{k} = e;
  end
  h3m_r.hmm{j} = hmm;
end

% ---- options (vbhem_h3m_cluster.m:141-210)
o = struct();
o.alpha0 = in.opt_alpha0;  o.eta0 = in.opt_eta0;  o.epsilon0 = in.opt_epsilon0;
o.lambda0 = in.opt_lambda0;  o.v0 = in.opt_v0;  o.W0 = in.opt_W0;  o.m0 = in.opt_m0(:);
o.Nv = in.opt_Nv;  o.tau = in.opt_tau;  o.max_iter = in.opt_max_iter;  o.minDiff = in.opt_minDiff;
o.emit.covar_type = 'full';
o.verbose = 0;  o.calc_LLderiv = 0;  o.useMEX = 0;  o.S = S;  o.K = K;

out = vbhem_h3m_c_step_fc(h3m_r, h3m_b, o);

g = struct();
g.LogLs = out.LogLs;
g.Z = out.Z;  g.L_elbo = out.L_elbo;  g.Nj = out.Nj;
g.alpha = out.alpha;
g.eta = zeros(K, S);  g.epsilon = zeros(K, S, S);  g.m = zeros(K, S, d);  g.W = zeros(K, S, d, d);
g.v = zeros(K, S);  g.lambda = zeros(K, S);
for j = 1:K
  vp = out.hmm{j}.varpar;                                      % form_outputH3M.m:15-28
  g.eta(j, :) = vp.alpha(:)';
  g.epsilon(j, :, :) = vp.epsilon;
  g.v(j, :) = vp.v(:)';  g.lambda(j, :) = vp.beta(:)';
  for k = 1:S
    g.m(j, k, :) = vp.m(:, k);
    g.W(j, k, :, :) = vp.W(:, :, k);
  end
end
save('-v7', fullfile(here, 'matlab_stage_golden.mat'), '-struct', 'g');
fprintf('wrote %s (%d iterations, LL = %.10g)\n', fullfile(here, 'matlab_stage_golden.mat'), numel(g.LogLs), g.LogLs(end));
