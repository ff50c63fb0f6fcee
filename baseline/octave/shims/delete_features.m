function [ filter, features_info ] = delete_features( filter, features_info )
% Missing in the reference (called by matlab_code/map_management.m:7, not shipped).  Rule of the
% published 1-point-RANSAC EKF toolbox the reference derives from: a feature that has been predicted
% more than 5 times and matched in fewer than half of those predictions is removed from the state
% (matlab_code/delete_a_feature.m does the state/covariance surgery) and from features_info.
% Highest index first, so the positions of the remaining candidates stay valid.
for i = length(features_info):-1:1
    if (features_info(i).times_measured < 0.5*features_info(i).times_predicted) && (features_info(i).times_predicted > 5)
        [ x_k_k, p_k_k ] = delete_a_feature( get_x_k_k(filter), get_p_k_k(filter), i, features_info );
        filter = set_x_k_k( filter, x_k_k );
        filter = set_p_k_k( filter, p_k_k );
        features_info = [ features_info(1:i-1) features_info(i+1:end) ];
    end
end
