% GPU drop-in for matlab_code/map_management.m (delete_features -> measured -> update_features_info ->
% inversedepth_2_cartesian -> initialize_features in ONE device call, ekfslam_map_management).  `im` is not an
% image here: the reference's corner search (initialize_a_feature.m:22-57) needs the Computer Vision Toolbox,
% so `im` carries the corner detections, a 3xK matrix [u; v; descriptor] (integer pixels inside the excluded
% band of initialize_a_feature.m:8).
function [ filter, features_info ] = map_management( filter, features_info, cam, im, min_number_of_features_in_image, step )
[ filter, features_info ] = ekfslam_mex( 'map_management', filter, features_info, cam, im, min_number_of_features_in_image, step );
