function [hmms, Ls] = vbhmm_learn_batch_b200(datas, K, vbopt, inits)
% VBHMM_LEARN_BATCH_B200  learn one K-state VB-HMM per subject on a B200 GPU, all subjects (and all random
% initialisations) in ONE device-resident loop (calls vbhmm_em_batch_mex).
%
%   [hmms, Ls] = vbhmm_learn_batch_b200(datas, K, vbopt, inits)
%
% Counterpart of vbhmm_learn_batch (which loops vbhmm_learn -> vbhmm_em over the subjects) for a single K:
%   datas   {S x 1}, datas{s} = {n_s x 1} cell of [T x dim] matrices (as for vbhmm_learn_batch)
%   K       number of states
%   vbopt   alpha0, epsilon0, beta0, v0, mu0, W0, maxIter, minDiff (vbhmm_learn's names)
%   inits   {S x numTrials} cell of mix_t structs as vbhmm_init returns them (fields alpha, epsilon, beta, v, m, W):
%           one vbhmm_em run per entry; per subject the run with the highest lower bound is kept (vbhmm_learn.m:555)
% Returns hmms {S x 1} with the fields vbhmm_em builds (prior, trans, pdf, LL, M, N1, N, varpar) and Ls [S x 1].
%
% See INTEGRATION.md section 3c.

[S, nt] = size(inits);
R = S * nt;
dim = length(vbopt.mu0);
alpha = zeros(K,R); epsilon = zeros(K,K,R); beta = zeros(K,R); v = zeros(K,R); m = zeros(dim,K,R); W = zeros(dim,dim,K,R);
run_subject = zeros(R,1);
r = 0;
for s = 1:S
  for t = 1:nt
    r = r + 1;  mix = inits{s,t};  run_subject(r) = s;
    alpha(:,r) = mix.alpha(:);  epsilon(:,:,r) = mix.epsilon;  beta(:,r) = mix.beta(:);  v(:,r) = mix.v(:);
    m(:,:,r) = mix.m;  W(:,:,:,r) = mix.W;
  end
end
prior = [vbopt.alpha0 vbopt.epsilon0 vbopt.beta0 vbopt.v0];

[alpha, epsilon, beta, v, m, W, N1, N, M, LLs, info] = vbhmm_em_batch_mex(datas, run_subject, K, prior, vbopt.mu0(:), ...
    vbopt.W0(:), alpha, epsilon, beta, v, m, W, vbopt.maxIter, vbopt.minDiff);

hmms = cell(S,1);  Ls = -inf(S,1);
for r = 1:R
  s = run_subject(r);
  L = LLs(max(info(1,r),1), r);
  if L <= Ls(s), continue; end
  Ls(s) = L;
  hmm = struct;
  hmm.prior = alpha(:,r) / sum(alpha(:,r));                                   % vbhmm_em.m:430-432
  hmm.trans = bsxfun(@rdivide, epsilon(:,:,r), sum(epsilon(:,:,r),2));        % :434-439
  hmm.pdf = cell(K,1);
  for k = 1:K                                                                 % :391-404, :455-456
    if v(k,r) > dim+1, C = inv(W(:,:,k,r)) / (v(k,r)-dim-1); else, C = inv(W(:,:,k,r)) / v(k,r); end
    hmm.pdf{k}.mean = m(:,k,r)';  hmm.pdf{k}.cov = (C + C')/2;
  end
  hmm.LL = L;  hmm.M = M(:,:,r);  hmm.N1 = N1(:,r);  hmm.N = N(:,r);
  hmm.varpar = struct('alpha', alpha(:,r), 'epsilon', epsilon(:,:,r), 'beta', beta(:,r), 'v', v(:,r), ...
                      'm', m(:,:,r), 'W', W(:,:,:,r));
  hmms{s} = hmm;
end
end
