% CFS_FANUC -- drop-in for Lib/CFS_FANUC.m whose optimizer() runs on the GPU (libcfs_b200.so through cfs_mex).
%
%   self = CFS_FANUC(obs, sys_info, ROBOT);  self = self.optimizer();
%   then read self.u, self.x_, self.eval.cost_all, self.eval.e_u_all, self.iter_O, self.total_iter
% exactly as main_FANUC.m:150-162 and RRTstar_CFS.m:194-202 do.  Put this directory BEFORE Lib/ on the MATLAB path.
% 'grad' selects num_jac (class path, default) or derivest (script path of M16iB/main_CFS.m).
classdef CFS_FANUC
    properties
        obs cell
        sys_info struct
        nn
        ROBOT = 'M16iB'
        grad = 'num_jac'
        u
        x_
        eval
        iter_O = 1
        total_iter = 0
        status = []
    end
    methods
        function self = CFS_FANUC(obs, sys_info, varargin)
            self.obs = obs;
            self.sys_info = sys_info;
            self.nn = sys_info.H * sys_info.nu;
            if ~isempty(varargin), self.ROBOT = varargin{1}; end
            if numel(varargin) > 1, self.grad = varargin{2}; end
            self.x_ = sys_info.x_;
            self.u = zeros(self.nn, 1);
            self.eval = struct('cost_all', [], 'e_cost_all', [], 'e_u_all', []);
        end
        function self = optimizer(self)
            [u, x, cost, eu, it, st, qp] = cfs_mex('CFS', self.grad, self.ROBOT, self.obs, self.sys_info);
            if bitand(st(1), 255) == 2
                % the reference's quadprog returns [] here and CFS_FANUC.m:92 throws an index error
                error('cfs:infeasible', 'QP infeasible at outer iteration %d', it(1) + 1);
            end
            k = double(it(1));
            self.u = u(:, 1);
            self.x_ = x(:, 1);
            self.status = st(1);
            self.iter_O = k + 1;
            self.total_iter = qp;
            self.eval.cost_all = cost(1:k, 1)';
            self.eval.e_u_all = eu(1:k, 1)';
            c0 = self.sys_info.caug;            % get_cost(u = 0), CFS_FANUC.m:63
            prev = [c0, self.eval.cost_all(1:end-1)];
            self.eval.e_cost_all = abs(prev(1:k) - self.eval.cost_all);
        end
    end
end
