% GPU drop-in for matlab_code/rescue_hi_inliers.m: forwards to the MEX gateway over libekfslam.so.
function features_info = rescue_hi_inliers( filter, features_info, cam )
features_info = ekfslam_mex( 'rescue_hi_inliers', filter, features_info, cam );
