% GPU drop-in for matlab_code/add_features_inverse_depth.m (hinv + add_a_feature_covariance_inverse_depth).
function [ X_RES, P_RES, newFeature ] = add_features_inverse_depth( uvd, X, P, cam, std_pxl, initial_rho, std_rho )
[ X_RES, P_RES, newFeature ] = ekfslam_mex( 'add_features_inverse_depth', uvd, X, full(P), cam, std_pxl, initial_rho, std_rho );
