function run_ref_step(inputs_mat, outputs_mat, ref_dir)
% Runs the UNMODIFIED reference functions (diwakar-vsingh/EKF-SLAM, matlab_code/) over the filter
% steps stored in inputs_mat (written by export_inputs.py) and saves what they produce, plus the
% wall time per filter-step.  GNU Octave or MATLAB.
%
%   octave --no-gui --eval "run_ref_step('inputs.mat','outputs.mat','/root/reference/matlab_code')"
%
% One filter-step = matlab_code/mono_slam.m:56-74 without the image front-end:
%   update_features_info -> ekf_prediction -> search_IC_matches lines 4-10 (predict_camera_measurements,
%   calculate_derivatives, S_i) -> the GATING rule of matching.m:16,38 on supplied candidate pixels
%   -> ransac_hypotheses -> ekf_update_li_inliers -> rescue_hi_inliers -> ekf_update_hi_inliers.
% Only three files are not the reference's own: shims/quaternions.m and shims/dq3_by_dq1.m (the
% reference calls them but does not ship them) and shims/select_random_match.m (reads the stored
% uniform stream instead of rand(1)).  A frame without individually compatible matches is a
% pass-through (the reference would fail at select_random_match.m:16).
if nargin < 3, ref_dir = '/root/reference/matlab_code'; end
here = fileparts(mfilename('fullpath'));
addpath(ref_dir);
addpath(fullfile(here, 'shims'));   % after ref_dir: the shims take precedence
global EKFSLAM_U EKFSLAM_UI
D = load(inputs_mat);
B = double(D.B); N = double(D.N); T = double(D.T);
cam = initialize_cam();
chi2inv_2_95 = 5.9915;
X = zeros(size(D.x0, 1), size(D.x0, 2), T);          % [B, n, T]
Pdiag = zeros(size(X));
Plast = zeros(size(D.P0));                           % [B, n, n] after the last frame
li = zeros(B, N, T); hi = zeros(B, N, T); ic = zeros(B, N, T); nhyp = zeros(B, T);
wall = 0;
for b = 1:B
    x = squeeze(D.x0(b, :))';
    P = squeeze(D.P0(b, :, :));
    filter = ekf_filter(x, P, double(D.std_a), double(D.std_alpha), double(D.std_z), 'constant_velocity');
    features_info = [];
    for i = 1:N
        f.type = 'inversedepth'; f.yi = x(13 + 6*(i-1) + (1:6));
        if D.types(b, i) == 2, f.type = 'cartesian'; end
        f.individually_compatible = 0; f.low_innovation_inlier = 0; f.high_innovation_inlier = 0;
        f.times_predicted = 0; f.times_measured = 0;
        f.z = []; f.h = []; f.H = []; f.S = []; f.R = eye(2);
        f.state_size = 6; f.measurement_size = 2;
        if isempty(features_info), features_info = f; else, features_info(i) = f; end
    end
    for t = 1:T
        EKFSLAM_U = squeeze(D.U(b, t, :)); EKFSLAM_UI = 0;
        t0 = tic;
        features_info = update_features_info(features_info);
        [filter, features_info] = ekf_prediction(filter, features_info);
        features_info = predict_camera_measurements(get_x_k_km1(filter), cam, features_info);
        features_info = calculate_derivatives(get_x_k_km1(filter), cam, features_info);
        nic = 0;
        for i = 1:N
            if ~isempty(features_info(i).h)
                features_info(i).S = features_info(i).H * get_p_k_km1(filter) * features_info(i).H' + features_info(i).R;
                if D.has(t, b, i)
                    S = full(features_info(i).S);
                    zc = squeeze(D.zc(t, b, i, :));
                    nu = zc - features_info(i).h';
                    if all(eig(S) < 100) && (nu' * inv(S) * nu < chi2inv_2_95)
                        features_info(i).individually_compatible = 1;
                        features_info(i).z = zc;
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
            filter.x_k_k = filter.x_k_km1; filter.p_k_k = filter.p_k_km1;
        end
        wall = wall + toc(t0);
        X(b, :, t) = filter.x_k_k(:)';
        Pdiag(b, :, t) = diag(full(filter.p_k_k))';
        nhyp(b, t) = EKFSLAM_UI;
        for i = 1:N
            ic(b, i, t) = features_info(i).individually_compatible;
            li(b, i, t) = features_info(i).low_innovation_inlier;
            hi(b, i, t) = features_info(i).high_innovation_inlier;
        end
    end
    Plast(b, :, :) = full(filter.p_k_k);
end
steps_per_s = B * T / wall;
printf('REFERENCE %d filter-steps in %.3f s = %.3f filter-steps/s (nproc %d)\n', B * T, wall, steps_per_s, nproc());
save('-v7', outputs_mat, 'X', 'Pdiag', 'Plast', 'li', 'hi', 'ic', 'nhyp', 'wall', 'steps_per_s');
end
