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
% The per-frame sequence is ref_frame.m and the features_info constructor ref_features_info.m (shared
% with the mini interpreter of oracle/mref, which executes the same reference sources where no Octave
% exists).  Only three files shadow / complete the reference: shims/quaternions.m and
% shims/dq3_by_dq1.m (the reference calls them but does not ship them) and
% shims_octave/select_random_match.m (reads the stored uniform stream instead of rand(1); the
% interpreter feeds rand() itself and runs the reference's own file).
if nargin < 3, ref_dir = '/root/reference/matlab_code'; end
here = fileparts(mfilename('fullpath'));
addpath(ref_dir);
addpath(fullfile(here, 'shims'));   % after ref_dir: the shims take precedence
addpath(fullfile(here, 'shims_octave'));
global EKFSLAM_U EKFSLAM_UI
D = load(inputs_mat);
B = double(D.B); N = double(D.N); T = double(D.T);
cam = initialize_cam();
X = zeros(size(D.x0, 1), size(D.x0, 2), T);          % [B, n, T]
Pdiag = zeros(size(X));
Plast = zeros(size(D.P0));                           % [B, n, n] after the last frame
li = zeros(B, N, T); hi = zeros(B, N, T); ic = zeros(B, N, T); nhyp = zeros(B, T);
wall = 0;
for b = 1:B
    x = squeeze(D.x0(b, :))';
    P = squeeze(D.P0(b, :, :));
    filter = ekf_filter(x, P, double(D.std_a), double(D.std_alpha), double(D.std_z), 'constant_velocity');
    features_info = ref_features_info(D.types(b, :));
    for t = 1:T
        EKFSLAM_U = squeeze(D.U(b, t, :)); EKFSLAM_UI = 0;
        t0 = tic;
        [filter, features_info] = ref_frame(filter, features_info, cam, squeeze(D.zc(t, b, :, :))', squeeze(D.has(t, b, :))');
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
