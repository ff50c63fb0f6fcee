function [filter, features_info, nic] = ref_frame_nomm(filter, features_info, cam, zc, has)
% Harness (NOT reference code): as ref_frame.m but WITHOUT the update_features_info call (map_management.m:17 has
% already run it in the closed loop of mono_slam.m:53): ONE filter step = matlab_code/mono_slam.m:56-74 without the image
% front-end, calling only the reference's own functions:
%   update_features_info (map_management.m:17) -> ekf_prediction -> search_IC_matches.m:4-10
%   (predict_camera_measurements, calculate_derivatives, S_i) -> the GATING rule of
%   matching.m:16,38 applied to supplied candidate pixels zc (2xN, has(i) = candidate present)
%   -> ransac_hypotheses -> ekf_update_li_inliers -> rescue_hi_inliers -> ekf_update_hi_inliers.
% A frame without individually compatible matches is a pass-through (the reference would fail at
% select_random_match.m:16).
chi2inv_2_95 = 5.9915;
[filter, features_info] = ekf_prediction(filter, features_info);
features_info = predict_camera_measurements(get_x_k_km1(filter), cam, features_info);
features_info = calculate_derivatives(get_x_k_km1(filter), cam, features_info);
nic = 0;
for i = 1:length(features_info)
    if ~isempty(features_info(i).h)
        features_info(i).S = features_info(i).H * get_p_k_km1(filter) * features_info(i).H' + features_info(i).R;
        if has(i)
            S = full(features_info(i).S);
            zi = zc(:, i);
            nu = zi - features_info(i).h';
            if all(eig(S) < 100) && (nu' * inv(S) * nu < chi2inv_2_95)
                features_info(i).individually_compatible = 1;
                features_info(i).z = zi;
                nic = nic + 1;
            end
        end
    end
end
if nic > 0
    features_info = ransac_hypotheses(filter, features_info, cam);
    filter = ekf_update_li_inliers(filter, features_info);
    features_info = rescue_hi_inliers(filter, features_info, cam);
    filter = ekf_update_hi_inliers(filter, features_info);
else
    filter.x_k_k = filter.x_k_km1;
    filter.p_k_k = filter.p_k_km1;
end
