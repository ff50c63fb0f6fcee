% GPU drop-in for matlab_code/search_IC_matches.m.  The reference matches FAST/FREAK features in the
% image (Computer Vision Toolbox); here `im` is a 3 x N matrix [u; v; has_candidate] of candidate
% pixels that go through the same gating rule (matching.m:16,38) on the device.
function features_info = search_IC_matches( filter, features_info, cam, im )
features_info = ekfslam_mex( 'search_IC_matches', filter, features_info, cam, im );
