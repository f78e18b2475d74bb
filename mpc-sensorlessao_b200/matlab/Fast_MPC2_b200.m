classdef Fast_MPC2_b200
    % Fast_MPC2_b200  Drop-in for the reference's Fast_MPC2 value class (Fast_MPC/VAR_2/Fast_MPC2.m:1-145;
    % VAR_1/Fast_MPC2.m when A2 is []), backed by the B200 C-ABI library through fmpc_mex.
    %
    % Same 23-argument constructor, same property names, same solver method names.  README.md:548-570 runs
    % unmodified after  Fast_MPC2 -> Fast_MPC2_b200.  x0 / x0_pre / w / xf / x_init may carry one COLUMN PER
    % INSTANCE (n x nb, ...) to solve a whole batch in one call; the result is then N x nb.
    % The device handle is cached per problem (every constructor constant, compared with isequal), so constructing the
    % object every control step, as the reference's closed loop does, costs nothing on the GPU side;
    % Fast_MPC2_b200.clear_cache() destroys the cached handles.
    properties
        Q; R; S; q; r; Qf; qf; x_min; x_max; u_min; u_max; du_min; du_max; T; x0; x0_pre; u_prev; A1; A2; B; w
        x_final; x_init
        nu0 = []       % optional explicit dual start (inf_newton_solver.m:2 draws rand(); [] = MATLAB default stream)
        device = 0
        ramp_rows = 0          % 1: VAR_1 ramp-rate rows (VAR_1/fast_mpc_ineq_const.m:58-79); needs u_prev, du_min, du_max
        var1_literal_bug = 0   % 1 (VAR_1 only): second block row of C at columns n:3n+m-1 as VAR_1/fast_mpc_eq_const.m:34-37
    end
    methods
        function cs = Fast_MPC2_b200(Q,R,S,Qf,q,r,qf,xmin,xmax,umin,umax,dumin,dumax,T,x0,x0_pre,u_prev,A1,A2,B,w,xf,x_init)
            if nargin > 1
                cs.Q = Q; cs.R = R; cs.S = S; cs.Qf = Qf; cs.q = q; cs.r = r; cs.qf = qf;
                cs.x_min = xmin; cs.x_max = xmax; cs.u_min = umin; cs.u_max = umax; cs.du_min = dumin; cs.du_max = dumax;
                cs.T = T; cs.x0 = x0; cs.x0_pre = x0_pre; cs.u_prev = u_prev; cs.A1 = A1; cs.A2 = A2; cs.B = B; cs.w = w;
                cs.x_final = xf; cs.x_init = x_init;
            end
        end
        function z_init = initialize(obj)                       % fast_mpc_init.m:12-26
            [n, m] = size(obj.B);
            if ~isempty(obj.x_init)
                if size(obj.x_init,1) ~= obj.T*(n+m), error('Initialization size mismatch (T*(n+m))'); end
                z_init = obj.x_init;
            else
                z_init = repmat([(obj.u_min+obj.u_max)/2; (obj.x_min+obj.x_max)/2], obj.T, 1);
            end
        end
        function x_opt = mpc_fixed_log_newton(obj, nw, k)       % Fast_MPC2.m:124-130
            if isempty(nw), nw = 1000; end                      % inf_newton_solver.m:4-8: nw = [] => max_iter = 1000
            x_opt = obj.run(0, struct('kappa', k, 'niters', nw), 0, 0);
        end
        function [u0, status, iters] = mpc_step_resident(obj, nw, k, reset)
            % One step of a closed loop whose solver state stays on the GPU (fmpc_step_r): warm start = the previous
            % call's solution shifted one stage, x0_pre = [] => previous x0, u_prev = [] => previous U(:,0).  Only x0 goes
            % to the device and only U(:,0) (README.md:589) comes back.  reset = true starts a loop (cold start).
            if isempty(nw), nw = 1000; end
            h = obj.handle(size(obj.x0, 2));
            [u0, status, iters] = fmpc_mex('step_r', h, struct('kappa', k, 'niters', nw), logical(reset), obj.x0, obj.x0_pre, ...
                                           obj.u_prev, obj.w, obj.x_final, obj.nu0);
        end
        function x_opt = mpc_fixed_log(obj, k)                  % Fast_MPC2.m:116-123
            x_opt = obj.run(1, struct('kappa', k), 0, 0);
        end
        function x_opt = mpc_fixed_newton(obj, nw)              % Fast_MPC2.m:131-144
            x_opt = obj.run(2, struct('niters', nw), 0, 0);
        end
        function x_opt = mpc_solve_full(obj)                    % Fast_MPC2.m:100-115
            x_opt = obj.run(3, struct(), 0, 0);
        end
        function x_opt = mpc_solve_check(obj, k_min, k_max)     % Fast_MPC2.m:88-99
            x_opt = obj.run(4, struct(), k_min, k_max);
        end
    end
    methods (Access = private)
        function x_opt = run(obj, mode, params, k_min, k_max)
            h = obj.handle(size(obj.x0, 2));
            if mode == 0
                x_opt = fmpc_mex('step', h, params, obj.x0, obj.x0_pre, obj.u_prev, obj.w, obj.x_final, obj.x_init, obj.nu0);
            else
                x_opt = fmpc_mex('frontend', h, mode, params, k_min, k_max, obj.x0, obj.x0_pre, obj.u_prev, obj.w, ...
                                 obj.x_final, obj.x_init, obj.nu0);
            end
        end
        function h = handle(obj, nb)
            h = Fast_MPC2_b200.cache('get', obj, nb);
        end
    end
    methods (Static)
        function clear_cache()
            % destroys every cached device handle (fmpc_destroy); call when the loop is over or before `clear mex`
            Fast_MPC2_b200.cache('clear', [], 0);
        end
        function h = cache(cmd, obj, nb)
            % Handles are cached per problem.  The key is EVERY constructor constant the handle holds on the GPU (matrices,
            % linear costs, all bounds, T, flags, batch capacity, device), compared entry by entry with isequal -- never a
            % checksum, which two different problems can share.
            persistent entries
            h = [];
            if isempty(entries), entries = {}; end
            if strcmp(cmd, 'clear')
                for i = 1:numel(entries), fmpc_mex('destroy', entries{i}.h); end
                entries = {};
                return
            end
            [n, m] = size(obj.B);
            sys = struct('n', n, 'm', m, 'T', obj.T, 'var_order', 1 + ~isempty(obj.A2), 'ramp_rows', obj.ramp_rows, ...
                         'var1_literal_bug', obj.var1_literal_bug, ...
                         'A1', obj.A1, 'A2', obj.A2, 'B', obj.B, 'Q', obj.Q, 'R', obj.R, 'Qf', obj.Qf, 'q', obj.q, 'r', obj.r, ...
                         'qf', obj.qf, 'x_min', obj.x_min, 'x_max', obj.x_max, 'u_min', obj.u_min, 'u_max', obj.u_max, ...
                         'du_min', obj.du_min, 'du_max', obj.du_max);
            key = struct('sys', sys, 'nb', max(nb, 1), 'device', obj.device);
            for i = 1:numel(entries)
                if isequal(entries{i}.key, key), h = entries{i}.h; return, end
            end
            if numel(entries) >= 8                               % least recently created goes first
                fmpc_mex('destroy', entries{1}.h);
                entries(1) = [];
            end
            h = fmpc_mex('create', sys, max(nb, 1), obj.device);
            entries{end + 1} = struct('key', key, 'h', h);
        end
    end
end
