function q = quaternions(v, theta)
% Missing in the reference (called by matlab_code/v2q.m:15): unit quaternion of a rotation of
% theta about the unit axis v, scalar first -- [cos(theta/2); sin(theta/2)*v/||v||].
v = v(:);
q = [cos(theta/2); sin(theta/2) * v / norm(v)];
