% GPU drop-in for matlab_code/delete_a_feature.m.
function [ X_km1_km1_new, P_km1_km1_new ] = delete_a_feature( X_km1_km1, P_km1_km1, featToDelete, features_info )
[ X_km1_km1_new, P_km1_km1_new ] = ekfslam_mex( 'delete_a_feature', X_km1_km1, full(P_km1_km1), featToDelete, features_info );
