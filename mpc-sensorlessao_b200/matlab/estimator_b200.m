function ad_est = estimator_b200(A_s, b_s, Y_M)
% estimator_b200  Batched replacement of the estimator line of the reference's closed loop (README.md:478):
%   ad_est = lsqminnorm((A_s'*A_s), ((A_s)'*(Y_M - b_s)));
% A_s : npix x nmodes model matrix of model_approx.mat with the piston column removed (README.md:289-290)
% b_s : npix x 1 offset; Y_M : npix x nb measurement vectors (one per column) -> ad_est : nmodes x nb
persistent he key
k = [size(A_s), sum(A_s(:)), sum(b_s(:))];
if isempty(he) || ~isequal(key, k)
    if ~isempty(he), fmpc_mex('est_destroy', he); end
    he = fmpc_mex('est_create', A_s, b_s, max(size(Y_M, 2), 4096), 0); key = k;
end
ad_est = fmpc_mex('est_apply', he, Y_M);
end
