function [K_ Te_ w_ Xe_ error_] = bundle_euclid_gpu( K, Te, w, Xe, x, varargin )
% BUNDLE_EUCLID_GPU  drop-in for BUNDLE_EUCLID (toolbox/bundle/bundle_euclid.m): same inputs,
% same option strings ('fix_structure', 'fix_motion', 'fix_pivot', pivot, 'fix_calibration',
% 'fix_principal', 'visibility', visible, 'verbose'), same outputs; the whole LM loop runs on
% the GPU inside mex_bundle_euclid_gpu (libvlgba).  NOTE: written against the reference's
% interface but never executed here -- the build container has no MATLAB/Octave.
m = size(w, 2);
flags = [4 0 0 0];            % num_variableK fix_structure fix_motion verbose
visible = [];
pivot = [];
k = 1;
while k <= numel(varargin)
    switch lower(varargin{k})
        case 'fix_structure',   flags(2) = 1;
        case 'fix_motion',      flags(3) = 1;
        case 'fix_pivot',       pivot = double(varargin{k+1}); k = k + 1;
        case 'fix_calibration', flags(1) = 0;
        case 'fix_principal',   flags(1) = 1;
        case 'visibility',      visible = double(varargin{k+1}); k = k + 1;
        case 'verbose',         flags(4) = 1;
    end
    k = k + 1;
end
if ~isempty(pivot), pivot = reshape(pivot, 1, m); end
[K_ Te_ w_ Xe_ error_] = mex_bundle_euclid_gpu(K, Te, w, Xe, x, visible, pivot, flags);
end
