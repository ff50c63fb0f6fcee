% GPU drop-in for matlab_code/ekf_prediction.m: forwards to the MEX gateway over libekfslam.so.
function [ f, features_info ] = ekf_prediction( f, features_info )
[ f, features_info ] = ekfslam_mex( 'ekf_prediction', f, features_info );
