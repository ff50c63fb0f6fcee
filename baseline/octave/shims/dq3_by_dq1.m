function M = dq3_by_dq1(q)
% Missing in the reference (called by matlab_code/dfv_by_dxv.m:13 and func_Q.m:24): derivative of
% q3 = q1 (x) q2 with respect to q2 evaluated with q = q1, i.e. the left-multiplication matrix of q
% (from the product rule of matlab_code/qprod.m:8).
r = q(1); x = q(2); y = q(3); z = q(4);
M = [r -x -y -z;
     x  r -z  y;
     y  z  r -x;
     z -y  x  r];
