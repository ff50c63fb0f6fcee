function [ filter, features_info ] = initialize_features( step, cam, filter, features_info, num_features_to_initialize, im )
% Shadows matlab_code/initialize_features.m + initialize_a_feature.m, whose corner search needs
% MATLAB's Computer Vision Toolbox (detectFASTFeatures / extractFeatures, initialize_a_feature.m:29,51)
% and an image.  Here `im` is the synthetic detector's output: im.uv (2xK integer corner pixels inside
% the excluded band, one attempt each), im.tag (1xK descriptor stand-ins) and im.img (blank image for the
% patch copy of add_feature_to_info_vector.m:7).  Everything after the corner is found is the
% reference's own code: add_features_inverse_depth (hinv + covariance augmentation) and
% add_feature_to_info_vector, with the constants of initialize_a_feature.m:10-12 and the attempt cap of
% initialize_features.m:5.
max_attempts = 50;
attempts = 0;
initialized = 0;
initial_rho = 1;
std_rho = 1;
std_pxl = get_std_z(filter);
K = size(im.uv, 2);
while ( initialized < num_features_to_initialize ) && ( attempts < max_attempts ) && ( attempts < K )
    attempts = attempts + 1;
    uv = im.uv(:, attempts);
    [ X_RES, P_RES, newFeature ] = add_features_inverse_depth( uv, get_x_k_k(filter), get_p_k_k(filter), cam, std_pxl, initial_rho, std_rho );
    filter = set_x_k_k(filter, X_RES);
    filter = set_p_k_k(filter, P_RES);
    features_info = add_feature_to_info_vector( uv, im.img, X_RES, features_info, step, newFeature, im.tag(attempts) );
    initialized = initialized + 1;
end
