function [zi, position, num_IC_matches] = select_random_match(features_info)
% Shadows matlab_code/select_random_match.m for parity runs: identical except that rand(1)
% (line 12) is replaced by the next value of a stored uniform stream, so that the Octave run, the
% CPU oracle and the GPU read the SAME numbers.  Hypothesis i of a frame uses u(i).
global EKFSLAM_U EKFSLAM_UI
map_size = length(features_info);
individually_compatible = zeros(map_size, 1);
for i = 1:map_size
    if features_info(i).individually_compatible
        individually_compatible(i) = 1;
    end
end
EKFSLAM_UI = EKFSLAM_UI + 1;
u = EKFSLAM_U(EKFSLAM_UI);
random_match_position = floor(u * sum(individually_compatible)) + 1;
positions_individually_compatible = find(individually_compatible);
position = positions_individually_compatible(random_match_position);
zi = features_info(position).z;
num_IC_matches = sum(individually_compatible);
