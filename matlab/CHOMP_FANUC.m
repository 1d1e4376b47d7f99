% CHOMP_FANUC -- drop-in for Lib/CHOMP_FANUC.m (optimizer() on the GPU through cfs_mex).
% Same constructor and result properties as the reference class: self = CHOMP_FANUC(obs, sys_info, uu, ROBOT); self = self.optimizer();
% The reference's loop never meets EVAL's convergence test (CHOMP_FANUC never updates eval.x_ / eval.x_old), so both run exactly
% sys_info.MAX_O_ITER updates u <- u - alpha*3*(QQ*u + ff + 2000*dcostObs)  (Lib/CHOMP_FANUC.m:56-83).
classdef CHOMP_FANUC
    properties
        obs cell
        sys_info struct
        nn
        ROBOT = 'M16iB'
        u
        x_
        eval
        iter_O = 1
        total_iter = 0
        status = []
    end
    methods
        function self = CHOMP_FANUC(obs, sys_info, uu, varargin)
            self.obs = obs;
            self.sys_info = sys_info;
            self.nn = sys_info.H * sys_info.nu;
            if ~isempty(varargin), self.ROBOT = varargin{1}; end
            self.x_ = sys_info.x_;
            self.u = uu;
            self.eval = struct('cost_all', [], 'e_cost_all', [], 'e_u_all', []);
        end
        function self = optimizer(self)
            [u, x, cost, eu, it, st] = cfs_mex('solve', 'CHOMP', 'derivest', self.ROBOT, self.obs, self.sys_info, self.u);
            k = double(it(1));
            self.u = u(:, 1);
            self.x_ = x(:, 1);
            self.status = st(1);
            self.iter_O = k + 1;
            self.eval.cost_all = cost(1:k, 1)';
            self.eval.e_u_all = eu(1:k, 1)';
        end
    end
end
