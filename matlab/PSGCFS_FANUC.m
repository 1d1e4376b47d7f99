% PSGCFS_FANUC -- drop-in for Lib/PSGCFS_FANUC.m (optimizer() on the GPU through cfs_mex).
% The normrnd(0,0.1,[nn,1]) draws of PSGCFS_FANUC.m:109 are made HERE, one column per outer iteration, and handed to the GPU so
% that MATLAB's random stream is the one consumed.  Difference to the reference: all MAX_O_ITER columns are drawn up front,
% whereas PSGCFS_FANUC.m:109 draws only when a PSG step is actually taken (stop_inner, :136-142, can skip it), so the
% generator may advance further here; pass your own draws through the optional argument of optimizer(noise) to control that.
% The gateway refuses PSGCFS without noise (a zero-noise run would be a different, deterministic method).
classdef PSGCFS_FANUC
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
        function self = PSGCFS_FANUC(obs, sys_info, varargin)
            self.obs = obs;
            self.sys_info = sys_info;
            self.nn = sys_info.H * sys_info.nu;
            if ~isempty(varargin), self.ROBOT = varargin{1}; end
            self.x_ = sys_info.x_;
            self.u = zeros(self.nn, 1);
            self.eval = struct('cost_all', [], 'e_cost_all', [], 'e_u_all', []);
        end
        function self = optimizer(self, varargin)
            K = self.sys_info.MAX_O_ITER;
            if ~isempty(varargin)
                noise = varargin{1};                     % nn x MAX_O_ITER, the caller's normrnd(0,0.1,[nn,1]) columns
            else
                noise = zeros(self.nn, K);
                for k = 1:K
                    noise(:, k) = normrnd(0, 0.1, [self.nn, 1]);
                end
            end
            [u, x, cost, eu, it, st, qp] = cfs_mex('PSGCFS', 'num_jac', self.ROBOT, self.obs, self.sys_info, noise);
            if bitand(st(1), 255) == 2
                error('cfs:infeasible', 'projection QP infeasible at outer iteration %d', it(1) + 1);
            end
            k = double(it(1));
            self.u = u(:, 1);
            self.x_ = x(:, 1);
            self.status = st(1);
            self.iter_O = k + 1;
            self.total_iter = qp;
            self.eval.cost_all = cost(1:k, 1)';
            self.eval.e_u_all = eu(1:k, 1)';
        end
    end
end
