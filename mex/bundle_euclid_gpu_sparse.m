function [K_ Te_ w_ Xe_ error_] = bundle_euclid_gpu_sparse( K, Te, w, Xe, obs_xy, obs_pt, obs_cam, varargin )
% BUNDLE_EUCLID_GPU_SPARSE  BUNDLE_EUCLID (toolbox/bundle/bundle_euclid.m) on an observation list:
%   obs_xy (2 x nobs), obs_pt, obs_cam (1 x nobs, 1-based) replace x (3 x n x m) and 'visibility' (n x m).
% The list must be in the reference's traversal order (camera-major, points ascending within a camera):
%   [obs_pt, obs_cam] = find(visibility); obs_xy = [x(1, visibility ~= 0)'; x(2, visibility ~= 0)'] ... or see
%   dense_to_list below.  Options as BUNDLE_EUCLID except 'visibility': 'fix_structure', 'fix_motion',
%   'fix_pivot', pivot, 'fix_calibration', 'fix_principal', 'verbose'.
% NOTE: written against the reference's interface but never executed here -- no MATLAB/Octave in the
% build container; the mex entry itself is tested through the shim (tests/test_mex_dropin.py).
m = size(w, 2);
flags = [4 0 0 0];            % num_variableK fix_structure fix_motion verbose
pivot = [];
k = 1;
while k <= numel(varargin)
    switch lower(varargin{k})
        case 'fix_structure',   flags(2) = 1;
        case 'fix_motion',      flags(3) = 1;
        case 'fix_pivot',       pivot = double(varargin{k+1}); k = k + 1;
        case 'fix_calibration', flags(1) = 0;
        case 'fix_principal',   flags(1) = 1;
        case 'verbose',         flags(4) = 1;
    end
    k = k + 1;
end
if ~isempty(pivot), pivot = reshape(pivot, 1, m); end
[K_ Te_ w_ Xe_ error_] = mex_bundle_euclid_gpu_sparse(K, Te, w, Xe, double(obs_xy), double(obs_pt(:)'), double(obs_cam(:)'), pivot, flags);
end

function [obs_xy, obs_pt, obs_cam] = dense_to_list( x, visibility ) %#ok<DEFNU>
% the list BUNDLE_EUCLID's dense arguments stand for (bundle_euclid.m:50,81,102)
[obs_pt, obs_cam] = find(visibility ~= 0);            % column-major: ascending i + n*j
n = size(visibility, 1);
idx = obs_pt + n * (obs_cam - 1);
x1 = reshape(x(1, :, :), [], 1); x2 = reshape(x(2, :, :), [], 1);
obs_xy = [x1(idx)'; x2(idx)'];
end
