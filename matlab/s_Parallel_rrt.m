%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
% Parallel RRT (script fragment) -- drop-in for Lib/functions/s_Parallel_rrt.m.
%
% Same workspace contract as the reference script (reads obs, sys_info, goalxyz, region_g, region_s, sample_off; leaves self,
% self_, routeL, path_fail, path_length, id, iter_rrt), but the num_seed trees of one round are grown by ONE GPU call
% (cfs_mex('rrt', ...): one thread block per seed) instead of a parfor over MATLAB workers (s_Parallel_rrt.m:16-25), so
% num_seed no longer "depends on #cores of the computer": thousands of seeds cost about the same as six.
% rand(NRND, num_seed) is drawn here, column i is the stream of seed i (the parfor workers of the reference each had their
% own generator; reseed with rng(...) to replay a run).
% On a multi-GPU box run one MATLAB worker per GPU and bind it first:  cfs_mex('device', mod(labindex - 1, gpuDeviceCount));
%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%%
if ~exist('num_seed', 'var'), num_seed = 6; end        % s_Parallel_rrt.m:11
if ~exist('NRND', 'var'), NRND = 4096; end
self_ = {};
self = [];
path_fail = true(num_seed, 1);
iter_rrt = 0;
MAX_ITER_RRT = 400;                                      % Lib/RRT_FANUC.m:37
while all(path_fail)
    routeL = 1000 * ones(num_seed, 1);
    rnd = rand(NRND, num_seed);
    [routes, len, nn, fl, used, nodes, parent, tot] = cfs_mex('rrt', 'M200i', 'RRT', obs, sys_info, goalxyz, region_g, ...
        region_s, sample_off, rnd, 0.5, MAX_ITER_RRT);
    cap = MAX_ITER_RRT + 2;
    routes = reshape(routes, sys_info.nstate, cap, num_seed);
    nodes = reshape(nodes, sys_info.nstate, cap, num_seed);
    for i = 1:num_seed
        r = RRT_FANUC(obs, sys_info, goalxyz, region_g, region_s, sample_off, 'M200i', 'RRT');
        k = double(nn(i));
        r.fail = logical(fl(i)) || len(i) < 0;
        r.route = routes(:, 1:max(double(len(i)), 0), i);
        r.node_num = k;
        r.all_nodes = [double(parent(1:k, i))'; nodes(:, 1:k, i)];
        r.total_dis = tot(1:k, i)';
        r.rnd_used = double(used(i));
        self_{i} = r;
        path_fail(i) = r.fail;
        if ~r.fail
            routeL(i) = size(r.route, 2);
        end
    end
    iter_rrt = iter_rrt + 1;
end
%%
[path_length, id] = min(routeL);
self = self_{id};
