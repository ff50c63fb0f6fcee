function features_info = ref_features_info(types)
% Harness helper (NOT reference code): a features_info struct array with the fields the hot path
% reads (matlab_code/add_feature_to_info_vector.m:7-32) for a map whose feature types are given
% as 1 = 'inversedepth', 2 = 'cartesian' (0 = unused slot, skipped).
features_info = [];
k = 0;
for i = 1:length(types)
    if types(i) ~= 0
        k = k + 1;
        features_info(k).type = 'inversedepth';
        if types(i) == 2
            features_info(k).type = 'cartesian';
        end
        features_info(k).times_predicted = 0;
        features_info(k).times_measured = 0;
        features_info(k).individually_compatible = 0;
        features_info(k).low_innovation_inlier = 0;
        features_info(k).high_innovation_inlier = 0;
        features_info(k).z = [];
        features_info(k).h = [];
        features_info(k).H = [];
        features_info(k).S = [];
        features_info(k).state_size = 6;
        features_info(k).measurement_size = 2;
        features_info(k).R = eye(2);
    end
end
