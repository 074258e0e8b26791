function [Pp_ Xp_ error_] = bundle_projective_gpu( Pp, Xp, x, varargin )
% BUNDLE_PROJECTIVE_GPU  drop-in for BUNDLE_PROJECTIVE (toolbox/bundle/bundle_projective.m): same inputs,
% same option strings ('fix_structure', 'fix_motion', 'visibility', visible, 'verbose'), same outputs; the
% whole LM loop runs on the GPU inside mex_bundle_projective_gpu (libvlgba).  NOTE: written against the
% reference's interface but never executed here -- the build container has no MATLAB/Octave.
flags = [0 0 0];            % fix_structure fix_motion verbose
visible = [];
k = 1;
while k <= numel(varargin)
    switch lower(varargin{k})
        case 'fix_structure',   flags(1) = 1;
        case 'fix_motion',      flags(2) = 1;
        case 'visibility',      visible = double(varargin{k+1}); k = k + 1;
        case 'verbose',         flags(3) = 1;
    end
    k = k + 1;
end
[Pp_ Xp_ error_] = mex_bundle_projective_gpu(Pp, Xp, x, visible, flags);
end
