function ad = zernmodfit_b200(frames, N)
% zernmodfit_b200  Batched replacement of the README.md:78-93 frame loop around zernmodfit.m:
%   for j: ad_new = zernmodfit(r(is_in), theta(is_in), z(is_in), N); ad_acc(j,:) = ad_new(:,1)'
% frames : nL x nL x nFrames phase screens (values outside the unit pupil are ignored, may be NaN)
% ad     : nFrames x nmodes coefficient table (= ad_acc), modes ordered n = 0..N, m = -n:2:n (zernmodfit.m:195-198)
persistent hz key
nL = size(frames, 1);
k = [nL, N];
if isempty(hz) || ~isequal(key, k)
    if ~isempty(hz), fmpc_mex('zmf_destroy', hz); end
    hz = fmpc_mex('zmf_create', nL, N, max(size(frames, 3), 2048), 0); key = k;
end
ad = fmpc_mex('zmf_fit', hz, frames)';
end
