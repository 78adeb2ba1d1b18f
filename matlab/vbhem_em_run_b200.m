function [h3m_r, LogLs, iters, unstable] = vbhem_em_run_b200(h3m_r, h3m_b, vbhemopt)
% VBHEM_EM_RUN_B200  the whole EM loop of vbhem_h3m_c_step_fc on a B200 GPU (calls vbhem_em_run_mex).
%
%   [h3m_r, LogLs, iters, unstable] = vbhem_em_run_b200(h3m_r, h3m_b, vbhemopt)
%
% Drop-in for the `while 1` loop of vbhem_h3m_c_step_fc (its lines 115-433) for full- or diagonal-covariance
% emissions: E-step, lower bound (vbhemh3m_lb), relative-change convergence test, NaN guard and M-step
% (vbhem_mstep_component) run on the device; the centre hyper-parameters never leave it between iterations.
%   h3m_r      initial centres as vbhemhmm_init returns them: alpha [1 x Kr], hmm{j}.eta, hmm{j}.epsilon,
%              hmm{j}.emit{k}.{m, W, v, lambda}
%   h3m_b      base mixture of hmms_to_h3m_hem
%   vbhemopt   alpha0, eta0, epsilon0, lambda0, v0, m0, W0, tau, max_iter, minDiff, Nv
% Returns h3m_r with the hyper-parameters after the last M-step, the trace of lower bounds (h3m_r.LogLs of the
% reference), the number of M-steps taken and whether the NaN guard fired.  form_outputH3M / one more E-step
% (vbhem_estep_b200) give hat_Z and L_elbo afterwards.
%
% See INTEGRATION.md section 3a.

Kr   = h3m_r.K;
Kb   = h3m_b.K;
dim  = h3m_b.hmm{1}.emit{1}.nin;
full = strcmp(h3m_b.hmm{1}.emit{1}.covar_type, 'full');
maxN = max(cellfun(@(x) length(x.prior), h3m_b.hmm));
Sr   = cellfun(@(x) length(x.eta), h3m_r.hmm);
S    = max(Sr);
tilde_N_k = vbhemopt.Nv * Kb * h3m_b.omega(:);          % vbhem_h3m_c_step_fc.m:26-30

eta = zeros(S,Kr); eps = zeros(S,S,Kr); m = zeros(dim,S,Kr); v = zeros(S,Kr); lam = zeros(S,Kr);
if full, W = zeros(dim,dim,S,Kr); else, W = zeros(dim,S,Kr); end
for j = 1:Kr
  s = Sr(j);
  eta(1:s,j)     = h3m_r.hmm{j}.eta(:);
  eps(1:s,1:s,j) = h3m_r.hmm{j}.epsilon;
  for k = 1:s
    e = h3m_r.hmm{j}.emit{k};
    m(:,k,j) = e.m(:);  v(k,j) = e.v;  lam(k,j) = e.lambda;
    if full, W(:,:,k,j) = e.W; else, W(:,k,j) = e.W(:); end
  end
end
prior = [vbhemopt.alpha0 vbhemopt.eta0 vbhemopt.epsilon0 vbhemopt.lambda0 vbhemopt.v0];
% vbhemopt.ngpus (optional): shard the base HMMs over that many GPUs of the box, 0 = all of them
if isfield(vbhemopt, 'ngpus'), ng = {vbhemopt.ngpus}; else, ng = {}; end

[alpha, eta, eps, m, W, v, lam, LogLs, info] = vbhem_em_run_mex(h3m_b.hmm, maxN, tilde_N_k, h3m_r.alpha(:)', ...
    eta, eps, m, W, v, lam, Sr(:)', prior, vbhemopt.m0(:), vbhemopt.W0(:), vbhemopt.tau, vbhemopt.max_iter, ...
    vbhemopt.minDiff, ng{:});

h3m_r.alpha = alpha;
for j = 1:Kr
  s = Sr(j);
  h3m_r.hmm{j}.eta     = eta(1:s,j);
  h3m_r.hmm{j}.epsilon = eps(1:s,1:s,j);
  for k = 1:s
    h3m_r.hmm{j}.emit{k}.m      = m(:,k,j)';
    h3m_r.hmm{j}.emit{k}.v      = v(k,j);
    h3m_r.hmm{j}.emit{k}.lambda = lam(k,j);
    if full, h3m_r.hmm{j}.emit{k}.W = W(:,:,k,j); else, h3m_r.hmm{j}.emit{k}.W = W(:,k,j)'; end
  end
end
h3m_r.LogLs = LogLs;
h3m_r.LL    = LogLs(end);
iters    = info(1);
unstable = info(2);
end
