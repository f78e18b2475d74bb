classdef Fast_MPC2_VAR1_b200 < Fast_MPC2_b200
    % Fast_MPC2_VAR1_b200  Drop-in for the reference's VAR(1) class (Fast_MPC/VAR_1/Fast_MPC2.m:1-141): same
    % 21-argument constructor (no x0_pre, single A), same method names, and -- by default -- the same problem the
    % reference assembles: ramp-rate rows (VAR_1/fast_mpc_ineq_const.m:58-79) and the second block row of C at
    % columns n : 3n+m-1 (VAR_1/fast_mpc_eq_const.m:34-37).  Set  obj.var1_literal_bug = 0  for the corrected
    % placement m+1 : 2(n+m), obj.ramp_rows = 0  for box rows only.
    methods
        function cs = Fast_MPC2_VAR1_b200(Q,R,S,Qf,q,r,qf,xmin,xmax,umin,umax,dumin,dumax,T,x0,u_prev,A,B,w,xf,x_init)
            cs = cs@Fast_MPC2_b200(Q,R,S,Qf,q,r,qf,xmin,xmax,umin,umax,dumin,dumax,T,x0,[],u_prev,A,[],B,w,xf,x_init);
            cs.ramp_rows = 1;
            cs.var1_literal_bug = 1;
        end
    end
end
