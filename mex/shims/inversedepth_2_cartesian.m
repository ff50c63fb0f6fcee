% GPU drop-in for matlab_code/inversedepth_2_cartesian.m (at most one conversion per call, :49).
function [ filter, features_info ] = inversedepth_2_cartesian( filter, features_info )
[ filter, features_info ] = ekfslam_mex( 'inversedepth_2_cartesian', filter, features_info );
