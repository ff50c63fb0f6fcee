% GPU drop-in for matlab_code/ransac_hypotheses.m.  The reference draws one rand(1) per executed
% hypothesis (select_random_match.m:12); the device replays a pre-drawn stream, so this shim draws
% the 1000 uniforms the frozen loop bound allows (ransac_hypotheses.m:9,14).  NOTE: that consumes
% more of the global stream than the reference does — set the global EKFSLAM_U to a fixed vector
% for bit-parity experiments.
function features_info = ransac_hypotheses( filter, features_info, cam )
global EKFSLAM_U
if isempty( EKFSLAM_U )
    u = rand( 1000, 1 );
else
    u = EKFSLAM_U;
end
features_info = ekfslam_mex( 'ransac_hypotheses', filter, features_info, cam, u );
