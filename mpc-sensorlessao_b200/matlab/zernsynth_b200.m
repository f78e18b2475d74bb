function phase_cor = zernsynth_b200(ad_cor, nL, N)
% zernsynth_b200  Batched replacement of the DM phase synthesis loop of the reference (README.md:592-598):
%   for j1 = 1:nx, Zs_cor(j1,:,:) = ad_cor(j1) .* squeeze(Zs(j1,:,:)); end;  phase_cor = squeeze(sum(Zs_cor, 1));
% ad_cor : nmodes x nFrames coefficients (modes ordered n = 0..N, m = -n:2:n; prepend a 0 for the removed piston)
% phase_cor : nL x nL x nFrames, zero outside the unit pupil
persistent hz key
k = [nL, N];
if isempty(hz) || ~isequal(key, k)
    if ~isempty(hz), fmpc_mex('zmf_destroy', hz); end
    hz = fmpc_mex('zmf_create', nL, N, max(size(ad_cor, 2), 2048), 0); key = k;
end
phase_cor = fmpc_mex('zmf_synth', hz, ad_cor, nL);
end
