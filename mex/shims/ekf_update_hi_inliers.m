% GPU drop-in for matlab_code/ekf_update_hi_inliers.m: forwards to the MEX gateway over libekfslam.so.
function filter = ekf_update_hi_inliers( filter, features_info )
filter = ekfslam_mex( 'ekf_update_hi_inliers', filter, features_info );
