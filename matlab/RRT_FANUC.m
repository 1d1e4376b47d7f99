% RRT_FANUC -- drop-in for Lib/RRT_FANUC.m whose find_route() grows the tree on the GPU (libcfs_b200.so through cfs_mex).
%
%   self = RRT_FANUC(obs, sys_info, goal, region_g, region_s, sample_off, ROBOT, SOLVER);  self = self.find_route();
%   then read self.route, self.fail, self.node_num, self.all_nodes, self.total_dis exactly as RRTstar_CFS.m:76-100 and
%   Lib/functions/s_Parallel_rrt.m:16-22 do.  Put this directory BEFORE Lib/ on the MATLAB path.
%
% MATLAB's rand stream stays the one that is consumed: find_route() draws a block rand(NRND,1) HERE and the kernel consumes it
% in the reference's order (pp = rand; rand(nstate,1) when pp < bi -- Lib/RRT_FANUC.m:108,111).  Difference to the reference:
% the block is drawn up front, so the generator advances by NRND numbers per tree instead of by the numbers actually used
% (self.rnd_used); reseed (rng(...)) per call where a run must be replayed.
classdef RRT_FANUC
    properties
        obs cell
        sys_info struct
        goal
        region_g
        region_s
        sample_off
        ROBOT = 'M16iB'
        SOLVER = 'RRT*'
        route
        fail = 0
        node_num = 1
        all_nodes
        total_dis
        MAX_ITER = 400      % Lib/RRT_FANUC.m:37
        bi = 0.5            % Lib/RRT_FANUC.m:38
        NRND = 4096         % uniform numbers handed to one tree (a 400-node tree uses ~1 + 5*bi per sample)
        rnd_used = 0
    end
    methods
        function self = RRT_FANUC(obs, sys_info, goal, region_g, region_s, sample_off, varargin)
            self.obs = obs;
            self.sys_info = sys_info;
            self.goal = goal;
            self.region_g = region_g;
            self.region_s = region_s;
            self.sample_off = sample_off;
            if ~isempty(varargin)
                self.ROBOT = varargin{1};
                self.SOLVER = varargin{2};
            end
        end
        function self = find_route(self)
            rnd = rand(self.NRND, 1);
            [routes, len, nn, fl, used, nodes, parent, tot] = cfs_mex('rrt', self.ROBOT, self.SOLVER, self.obs, self.sys_info, ...
                self.goal, self.region_g, self.region_s, self.sample_off, rnd, self.bi, self.MAX_ITER);
            if len(1) < 0
                error('cfs:rrt', 'RRT_FANUC.find_route: the random block (%d numbers) ran dry; raise NRND', self.NRND);
            end
            ns = self.sys_info.nstate;
            cap = self.MAX_ITER + 2;
            routes = reshape(routes, ns, cap);
            nodes = reshape(nodes, ns, cap);
            k = double(nn(1));
            self.route = routes(:, 1:double(len(1)));
            self.fail = logical(fl(1));
            self.node_num = k;
            self.all_nodes = [double(parent(1:k))'; nodes(:, 1:k)];     % row 1: parent (-1 for the root), Lib/RRT_FANUC.m:66
            self.total_dis = tot(1:k)';
            self.rnd_used = double(used(1));
        end
    end
end
