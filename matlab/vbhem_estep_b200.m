function [L_elbo, hat_Z, Z_Ni, Nj, estats, lb] = vbhem_estep_b200(h3m_b, h3m_r, vbhemopt, tilde_N_k, logOmegaTilde)
% VBHEM_ESTEP_B200  fused VBHEM-H3M E-step on a B200 GPU (calls vbhem_estep_fused_mex).
%
%   [L_elbo, hat_Z, Z_Ni, Nj, estats, lb] = vbhem_estep_b200(h3m_b, h3m_r, vbhemopt, tilde_N_k, logOmegaTilde)
%
% Drop-in for the E-step block of vbhem_h3m_c_step_fc (its lines 168-283) plus the statistics
% reduction that vbhem_mstep_component otherwise runs per centre (vbhem_compute_Statistics):
%   h3m_r must already carry the pre-pass fields of the current iteration
%   (logATilde, logPiTilde, emit{k}.logLambdaTildePlusDdivlamda), exactly as for vbhem_hmm_bwd_fwd_mex.
% Outputs
%   L_elbo, hat_Z, Z_Ni [Kb x Kr], Nj [1 x Kr]     as computed in vbhem_h3m_c_step_fc
%   estats{j}  struct with Nj_rho1, Nj_rho2rho, Nj_rho, y_bar, S_plus_C of centre j
%              (what vbhem_compute_Statistics returns), ready for the M-step formulas
%   lb = [Lt1 Lt7]                                  the two N-sized terms of vbhemh3m_lb
%
% See INTEGRATION.md for the one-branch change in vbhem_h3m_c_step_fc.m that uses this function.

Kr    = h3m_r.K;
dim   = h3m_b.hmm{1}.emit{1}.nin;
T     = vbhemopt.tau;
maxN  = max(cellfun(@(x) length(x.prior), h3m_b.hmm));
maxN2 = max(cellfun(@(x) length(x.eta),   h3m_r.hmm));
% vbhemopt.ngpus (optional): shard the base HMMs over that many GPUs of the box, 0 = all of them
if isfield(vbhemopt, 'ngpus'), ng = {vbhemopt.ngpus}; else, ng = {}; end

switch h3m_b.hmm{1}.emit{1}.covar_type
  case 'diag'
    [L_elbo, hat_Z, Nj, R1, R2, R, Y, SC, lb] = vbhem_estep_fused_mex( ...
        h3m_b.hmm, h3m_r.hmm, T, maxN, maxN2, tilde_N_k(:), logOmegaTilde(:)', ng{:});
  case 'full'
    ldR = cell(1,Kr);  icR = cell(1,Kr);
    for j = 1:Kr
      S = length(h3m_r.hmm{j}.eta);
      ldR{j} = zeros(1,S);  icR{j} = zeros(dim,dim,S);
      for n = 1:S
        ldR{j}(n)     = h3m_r.hmm{j}.emit{n}.logLambdaTildePlusDdivlamda;
        icR{j}(:,:,n) = h3m_r.hmm{j}.emit{n}.v .* h3m_r.hmm{j}.emit{n}.W;
      end
    end
    [L_elbo, hat_Z, Nj, R1, R2, R, Y, SC, lb] = vbhem_estep_fused_mex( ...
        h3m_b.hmm, h3m_r.hmm, T, maxN, maxN2, tilde_N_k(:), logOmegaTilde(:)', ldR, icR, ng{:});
  otherwise
    error('not supported')
end

Z_Ni = hat_Z .* (tilde_N_k(:) * ones(1,Kr));

estats = cell(Kr,1);
for j = 1:Kr
  S = length(h3m_r.hmm{j}.eta);
  estats{j}.Nj_rho1    = R1(1:S, j);
  estats{j}.Nj_rho2rho = R2(1:S, 1:S, j);
  estats{j}.Nj_rho     = R(1:S, j);
  estats{j}.y_bar      = reshape(Y(1:S, :, j), [S dim]);
  if ndims(SC) == 4 || (dim == 1 && strcmp(h3m_b.hmm{1}.emit{1}.covar_type, 'full'))
    estats{j}.S_plus_C = reshape(SC(1:S, :, :, j), [S dim dim]);
  else
    estats{j}.S_plus_C = reshape(SC(1:S, :, j), [S dim]);
  end
end
end
